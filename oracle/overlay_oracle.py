"""Oracle (TEST INFRASTRUCTURE) for the overlay compositions that consume the path's arrays
(SURVEY.md 8f3): a numpy restatement of the array arithmetic in
``/root/reference/Vent_Analysis.py`` -- ``exportDICOM`` (:387-391), ``cropToData`` (:430-456),
``screenShot`` (:458-494, up to the array handed to PIL) and the GUI's ``colorBinary``
(:628-634).  PINNED: ``tests/test_oracle_golden.py`` checks it against
``tests/golden/overlay_small_b.npz``, arrays produced by the reference's own methods
(``tests/golden/make_golden_overlay.py``).
"""
from __future__ import annotations

import numpy as np


def normalize(x):
    """Vent_Analysis.py:233-237 / :460-464 (dtype follows numpy: float32 stays float32, integer
    input divides in float64)."""
    if (np.max(x) - np.min(x)) == 0:
        return x
    return (x - np.min(x)) / (np.max(x) - np.min(x))


def export_rgb(N4, defect):
    """``exportDICOM`` :387-391 -> uint8 [r, c, s, 3]."""
    BW = (normalize(np.abs(N4)) * (2 ** 8 - 1)).astype("uint8")
    RGB = np.zeros(N4.shape + (3,), dtype=np.uint8)
    RGB[..., 0] = BW * (defect == 0) + 255 * (defect == 1)
    RGB[..., 1] = BW * (defect == 0)
    RGB[..., 2] = BW * (defect == 0)
    return RGB


def color_binary(A, B):
    """GUI ``colorBinary`` :628-634 on a volume (the reference feeds it the 1 x slices montages of
    ``abs(A)`` and ``abs(B)``, :642-643, so the normalisation runs over the whole volume), followed by
    ``arrayToImage``'s ``astype(np.uint8)`` (:621) -> uint8 [r, c, s, 3]."""
    A = normalize(np.abs(A))
    B = np.abs(B)
    new = np.zeros(A.shape + (3,))
    new[..., 0] = A * (B == 0) + B
    new[..., 1] = A * (B == 0)
    new[..., 2] = A * (B == 0)
    return (new * 255).astype(np.uint8)


def crop_box(A, border=0, border_slices=False):
    """``cropToData`` :430-456 -> (r0, r1, c0, c1, s0, s1).  Index 0 never counts as "has signal"
    (``np.multiply(flag, range)`` is 0 there); raises IndexError like the reference when an axis has
    no other index with signal."""
    A = np.asarray(A)
    sl = [x for x in range(1, A.shape[2]) if np.sum(A[:, :, x]) > 0]
    ro = [x for x in range(1, A.shape[0]) if np.sum(A[x, :, :]) > 0]
    co = [x for x in range(1, A.shape[1]) if np.sum(A[:, x, :]) > 0]
    sb = border if border_slices else 0
    return (max(ro[0] - border, 0), min(ro[-1] + border + 1, A.shape[0]),
            max(co[0] - border, 0), min(co[-1] + border + 1, A.shape[1]),
            max(sl[0] - sb, 0), min(sl[-1] + sb + 1, A.shape[2]))


def montage(images, grid_shape):
    """``skimage.util.montage(images, grid_shape=..., padding_width=0, fill=0)`` for equally sized
    2-D images: image k goes to grid cell (k // ncols, k % ncols)."""
    nr, nc = grid_shape
    h, w = images[0].shape
    out = np.zeros((nr * h, nc * w), dtype=np.result_type(*images))
    for k, im in enumerate(images):
        i, j = divmod(k, nc)
        out[i * h:(i + 1) * h, j * w:(j + 1) * w] = im
    return out


def screenshot_image(HPvent, proton, N4, mask, mask_border, defect, CI, parula):
    """``screenShot`` :466-494 up to ``np.uint8(IMAGE*255)`` -> uint8 [7 R, S C, 3] (R, C, S: the box of
    ``cropToData(mask, border=5)``).  Raises IndexError like the reference when a CI value reaches
    40 mm (``parula[int(CI*64/40)]`` with a 64-row table)."""
    r0, r1, c0, c1, s0, s1 = crop_box(mask, border=5)
    ix = np.ix_(range(r0, r1), range(c0, c1), range(s0, s1))
    blank = np.zeros_like(HPvent[ix])
    pro = normalize(proton[ix])
    HP = normalize(HPvent[ix])
    n4 = normalize(N4[ix])
    border = normalize(mask_border[ix]) > 0
    dA = defect[ix] > 0
    ci = CI[ix]
    idx = (ci * 64 / 40).astype(np.int64)            # int() truncates; CI >= 0
    if idx.max(initial=0) >= parula.shape[0]:
        raise IndexError(f"index {idx.max()} is out of bounds for axis 0 with size {parula.shape[0]}")
    col = [parula[idx, ch] for ch in range(3)]
    chans = []
    for ch in range(3):
        chans.append(np.concatenate((blank, blank, pro, HP, n4 * (~border) + (0 if ch == 0 else 1) * border,
                                     n4 * (~dA) + dA if ch == 0 else n4 * (~dA),
                                     n4 * (ci == 0) + col[ch] * (ci > 0)), axis=2))
    S = n4.shape[2]
    mont = [montage([c3[:, :, k] for k in range(c3.shape[2])], (7, S)) for c3 in chans]
    return np.uint8(np.stack(mont, axis=2) * 255)
