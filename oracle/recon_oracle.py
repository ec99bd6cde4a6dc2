"""Oracle (TEST INFRASTRUCTURE) for the k-space reconstruction of ``process_RAW``
(``/root/reference/Vent_Analysis.py:536-540``): numpy restatement.  PINNED:
``tests/test_oracle_golden.py`` checks it against ``tests/golden/recon_small.npz``, produced by
executing those reference source lines themselves (``tests/golden/make_golden_recon.py``)."""
import numpy as np


def reconstruct(raw_K):
    """complex k-space ``[rows, cols, slices]`` -> complex128 image ``[cols, rows, slices]``."""
    raw_K = np.asarray(raw_K)
    img = np.zeros(raw_K.shape).astype(np.complex128)                                   # :536
    for k in range(raw_K.shape[2]):                                                     # :537
        img[:, :, k] = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(raw_K[:, :, k])))    # :538
    return np.transpose(img, (1, 0, 2))[:, ::-1, :]                                     # :539
