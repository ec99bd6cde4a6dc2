"""ctypes front end of ``n4_oracle.c`` (TEST INFRASTRUCTURE, **parity unpinned** -- see
the header of the C file).  Restates ``Vent_Analysis.N4_bias_correction``
(Vent_Analysis.py:316-334) = SimpleITK 2.3.1 ``N4BiasFieldCorrectionImageFilter()``
with all defaults."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _build

DEFAULTS = dict(max_iter=(50, 50, 50, 50), conv_thr=0.001, fwhm=0.15, wiener=0.01, nbins=200,
                ncp=(4, 4, 4), mask_label=1)

_vp = ctypes.c_void_p


def _p(a):
    return a.ctypes.data_as(_vp)


def n4(image, mask, **kw):
    """Returns dict(out f32, logbias f32, iters int32[levels], conv f32[levels, maxit],
    lattice f32[ncp...], )."""
    o = dict(DEFAULTS)
    o.update(kw)
    lib = _build.load()
    img = np.ascontiguousarray(image, dtype=np.float32)              # Vent_Analysis.py:322
    m = np.ascontiguousarray(np.asarray(mask).astype(np.float32).astype(np.uint8))  # :323, :327
    n0, n1, n2 = img.shape
    mi = np.asarray(o["max_iter"], dtype=np.int32)
    ncp = np.asarray(o["ncp"], dtype=np.int32)
    nl = len(mi)
    out = np.empty_like(img)
    lb = np.empty_like(img)
    iters = np.zeros(nl, dtype=np.int32)
    stride = int(mi.max())
    conv = np.empty((nl, stride), dtype=np.float32)
    fin = np.zeros(3, dtype=np.int32)
    cap = 1
    for a in ncp:
        c = int(a)
        for _ in range(nl - 1):
            c = 2 * c - 3
        cap *= c
    lat = np.zeros(cap, dtype=np.float32)
    fn = lib.n4_oracle
    fn.restype = ctypes.c_int
    rc = fn(_p(img), _p(m), ctypes.c_int(n0), ctypes.c_int(n1), ctypes.c_int(n2),
            ctypes.c_int(int(o["mask_label"])), ctypes.c_int(nl), _p(mi),
            ctypes.c_float(o["conv_thr"]), ctypes.c_float(o["fwhm"]), ctypes.c_float(o["wiener"]),
            ctypes.c_int(int(o["nbins"])), _p(ncp), _p(out), _p(lb), _p(iters), _p(conv),
            ctypes.c_int(stride), _p(lat), _p(fin))
    if rc:
        raise ValueError(f"n4_oracle rc={rc}")
    return dict(out=out, logbias=lb, iters=iters, conv=conv,
                lattice=lat[:int(np.prod(fin))].reshape(tuple(int(x) for x in fin)).copy())


def sharpen(U, m, nbins=200, fwhm=0.15, wiener=0.01):
    lib = _build.load()
    U = np.ascontiguousarray(U, dtype=np.float32)
    m = np.ascontiguousarray(m, dtype=np.uint8)
    S = np.zeros_like(U)
    E = np.empty(nbins, dtype=np.float32)
    H = np.empty(nbins, dtype=np.float32)
    rng = np.empty(2, dtype=np.float32)
    lib.n4_sharpen(_p(U), _p(m), ctypes.c_int64(U.size), ctypes.c_int(nbins), ctypes.c_float(fwhm),
                   ctypes.c_float(wiener), _p(S), _p(E), _p(H), _p(rng))
    return dict(S=S, E=E, H=H, bmin=rng[0], bmax=rng[1])


def fit(resid, m, ncp):
    lib = _build.load()
    r = np.ascontiguousarray(resid, dtype=np.float32)
    m = np.ascontiguousarray(m, dtype=np.uint8)
    ncp = np.asarray(ncp, dtype=np.int32)
    phi = np.empty(tuple(int(x) for x in ncp), dtype=np.float32)
    lib.n4_fit(_p(r), _p(m), *(ctypes.c_int(int(x)) for x in r.shape), _p(ncp), _p(phi))
    return phi


def evaluate(lattice, shape):
    lib = _build.load()
    lat = np.ascontiguousarray(lattice, dtype=np.float32)
    ncp = np.asarray(lat.shape, dtype=np.int32)
    f = np.empty(tuple(int(x) for x in shape), dtype=np.float32)
    lib.n4_eval(_p(lat), _p(ncp), *(ctypes.c_int(int(x)) for x in shape), _p(f))
    return f


def refine(lattice):
    lib = _build.load()
    lat = np.ascontiguousarray(lattice, dtype=np.float32)
    ncp = np.asarray(lat.shape, dtype=np.int32)
    out = np.empty(tuple(2 * int(x) - 3 for x in ncp), dtype=np.float32)
    nn = np.zeros(3, dtype=np.int32)
    lib.n4_refine(_p(lat), _p(ncp), _p(out), _p(nn))
    return out


def convergence(f1, f2, m):
    lib = _build.load()
    f1 = np.ascontiguousarray(f1, dtype=np.float32)
    f2 = np.ascontiguousarray(f2, dtype=np.float32)
    m = np.ascontiguousarray(m, dtype=np.uint8)
    fn = lib.n4_convergence
    fn.restype = ctypes.c_float
    return float(fn(_p(f1), _p(f2), _p(m), ctypes.c_int64(f1.size)))
