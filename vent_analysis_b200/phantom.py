"""Seeded synthetic hyperpolarised-xenon ventilation phantoms.

The reference ships no input data and no phantom generator, so the benchmark,
the parity tests and the golden fixtures all draw their inputs from here.
Everything is a pure function of ``(shape, vox, seed)``.

Named configurations (BASELINE.json ``configs``):

=====  ================  ===================  ====
name   shape [r, c, s]   vox [mm]             seed
=====  ================  ===================  ====
C1     (128, 128, 14)    [1.5, 1.5, 10.0]     0
C2     (128, 128, 14)    [2.0, 2.0, 11.5]     1
C3     (128, 128, 128)   [3.9, 3.9, 3.9]      2
C4     (256, 256, 64)    [1.5, 1.5, 3.0]      3
=====  ================  ===================  ====

C5 (the cohort) alternates C1 / C2 with seeds 0..n-1.
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    "C1": ((128, 128, 14), [1.5, 1.5, 10.0], 0),
    "C2": ((128, 128, 14), [2.0, 2.0, 11.5], 1),
    "C3": ((128, 128, 128), [3.9, 3.9, 3.9], 2),
    "C4": ((256, 256, 64), [1.5, 1.5, 3.0], 3),
}


def phantom(shape, vox, seed):
    """Return ``(HPvent float32 [r,c,s], mask uint8 [r,c,s])``.

    Two ellipsoidal lungs, 8 % tissue texture, 2-6 ellipsoidal ventilation
    defects (signal x0.10-0.35), a smooth multiplicative coil bias
    ``exp(a*x + b*y^2 + c*z)`` and Rician-like noise.  ``vox`` only steers the
    through-plane size of the defects so that anisotropic stacks get defects
    that span a physically plausible number of slices.
    """
    rng = np.random.default_rng(seed)
    n0, n1, n2 = (int(v) for v in shape)
    vox = np.asarray(vox, dtype=np.float64)
    r = np.arange(n0, dtype=np.float64)[:, None, None]
    c = np.arange(n1, dtype=np.float64)[None, :, None]
    s = np.arange(n2, dtype=np.float64)[None, None, :]

    mask = np.zeros((n0, n1, n2), dtype=bool)
    for cc in (0.27 * n1, 0.73 * n1):
        e = ((r - 0.5 * (n0 - 1)) / (0.31 * n0)) ** 2 \
            + ((c - cc) / (0.14 * n1)) ** 2 \
            + ((s - 0.5 * (n2 - 1)) / max(0.43 * n2, 0.75)) ** 2
        mask |= e <= 1.0

    tissue = 1000.0 * mask * (1.0 + 0.08 * rng.standard_normal((n0, n1, n2)))

    ndef = int(rng.integers(2, 7))
    idx = np.argwhere(mask)
    aniso = float(vox[2] / min(vox[0], vox[1]))
    for _ in range(ndef):
        ctr = idx[int(rng.integers(0, len(idx)))]
        rad = rng.uniform(3.0, 12.0, size=3) * min(n0, n1) / 128.0
        rad[2] = max(rad[2] / aniso, 0.8)
        e = ((r - ctr[0]) / rad[0]) ** 2 + ((c - ctr[1]) / rad[1]) ** 2 \
            + ((s - ctr[2]) / rad[2]) ** 2
        tissue = np.where(e <= 1.0, tissue * rng.uniform(0.10, 0.35), tissue)

    a, b, g = rng.uniform(-0.4, 0.4, size=3)
    xr = np.linspace(-1.0, 1.0, n0)[:, None, None] if n0 > 1 else np.zeros((1, 1, 1))
    yc = np.linspace(-1.0, 1.0, n1)[None, :, None] if n1 > 1 else np.zeros((1, 1, 1))
    zs = np.linspace(-1.0, 1.0, n2)[None, None, :] if n2 > 1 else np.zeros((1, 1, 1))
    bias = np.exp(a * xr + b * yc ** 2 + g * zs)

    img = np.abs(tissue * bias + 30.0 * rng.standard_normal((n0, n1, n2)))
    return img.astype(np.float32), mask.astype(np.uint8)


def config(name):
    """``(HPvent, mask, vox)`` of one of the named configurations."""
    shape, vox, seed = CONFIGS[name]
    img, mask = phantom(shape, vox, seed)
    return img, mask, list(vox)


def cohort_member(i):
    """Subject ``i`` of the C5 cohort: C1 / C2 geometry alternating, seed ``i``."""
    shape, vox, _ = CONFIGS["C1" if i % 2 == 0 else "C2"]
    img, mask = phantom(shape, vox, i)
    return img, mask, list(vox)
