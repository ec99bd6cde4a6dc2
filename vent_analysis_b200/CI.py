"""Drop-in for the reference's ``CI`` module (``/root/reference/CI.py``): same function
names and signatures, the sphere-growing search runs in the sm_100a kernels of
``libventb200`` (``csrc/ci_kernels.cuh``) instead of a thread pool of ``np.intersect1d`` calls.

``calculate_CI`` is bit-exact with ``CI.calculate_CI`` (float64 CI array), including the two
reference quirks that matter: neighbour coordinates are never bounds-checked (the sphere
lives in flat row-fastest index space, CI.py:65-68,:93) and duplicated table offsets count
once in the numerator but every time in the denominator (CI.py:96).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .sphere_table import get_sphere_pix


def getSpherePix(vox, radius):
    """CI.getSpherePix (CI.py:33-63): ``N x 4`` float64 ``[radius, dx, dy, dz]``."""
    return get_sphere_pix(vox, radius)


def px2vec(i, j, k, arrayShape):
    """CI.px2vec (CI.py:65-68)."""
    return i + (j - 1) * arrayShape[0] + (k - 1) * arrayShape[0] * arrayShape[1]


def vec2px(n, arrayShape):
    """CI.vec2px (CI.py:70-77)."""
    s = np.ceil(n / (arrayShape[0] * arrayShape[1]))
    n = n - (s - 1) * arrayShape[1] * arrayShape[0]
    c = np.ceil(n / arrayShape[0])
    r = n - (c - 1) * arrayShape[0]
    return int(r), int(c), int(s)


def getRadiiIndices(data):
    """CI.getRadiiIndices (CI.py:79-85): row indices where a new radius starts."""
    return np.where(np.diff(data[:, 0]) > 0)[0] + 1


def multi_which(A):
    """CI.multi_which (CI.py:10-30): indices of the non-zero entries, C order."""
    if np.isscalar(A):
        return np.where(A)[0]
    return np.argwhere(np.asarray(A) != 0)


def _run(defectArray, vox, Rmax=50, device=0, ctx=None):
    """(CIarray float64 in table units, t95 radius in table units or None)."""
    ctx = ctx if ctx is not None else _lib.context(device)
    d = np.asarray(defectArray)
    if d.ndim != 3:
        raise ValueError("defectArray must be a 3D array")
    d8 = np.ascontiguousarray(d != 0, dtype=np.uint8)
    if not d8.any():
        # reference: multi_which -> empty (0,) array -> defList[:,0] raises IndexError (CI.py:118)
        raise IndexError("too many indices for array: array is 1-dimensional, but 2 were indexed")
    vox = [float(v) for v in np.asarray(vox, dtype=np.float64)]
    table = get_sphere_pix(vox, Rmax)
    h = ctx.ci_table(table, d8.shape, key=(tuple(vox), int(Rmax), table.shape[0]))
    shell, stats, rc = ctx.ci_host(d8, [d8.shape], [h])
    if rc == _lib.VB_ERR_CI_RMAX:
        raise ValueError("--MAX RADIUS REACHED-- (sphere table exhausted, CI.py:101-103)")
    radii = ctx.ci_radii(h)
    shell = shell.reshape(d8.shape)
    sel = d8 != 0
    cv = np.zeros(d8.shape, dtype=np.float64)
    cv[sel] = radii[shell[sel]]
    t95 = int(stats["ci_t95"][0])
    return cv, (radii[t95] if t95 >= 0 else None)


def calculate_CV(defectArrayShape, activeVoxel, defVec, spherePx):
    """CI.calculate_CV (CI.py:87-105) for ONE voxel -> ``[i, j, k, radius]``.  ``defVec`` are
    px2vec indices of the defect voxels; the search itself runs on the GPU."""
    shape = tuple(int(s) for s in defectArrayShape)
    n0, n1 = shape[0], shape[1]
    f = np.asarray(defVec, dtype=np.int64) + n0 + n0 * n1          # undo px2vec's (j-1), (k-1)
    d = np.zeros(shape, dtype=np.uint8)
    ok = (f >= 0) & (f < d.size)
    f = f[ok]
    d[f % n0, (f // n0) % n1, f // (n0 * n1)] = 1
    ctx = _lib.context(0)
    h = ctx.ci_table(spherePx, shape)
    shell, _, rc = ctx.ci_host(d, [shape], [h])
    a = [int(x) for x in activeVoxel[:3]]
    t = int(shell.reshape(shape)[a[0], a[1], a[2]])
    if rc == _lib.VB_ERR_CI_RMAX and t >= 0xFFFE:
        raise ValueError("--MAX RADIUS REACHED--")
    if t == _lib.CI_OFF:
        raise ValueError("activeVoxel is not a defect voxel")
    return np.append(np.asarray(activeVoxel), ctx.ci_radii(h)[t])


def calculate_CI(defectArray, vox=[1, 1, 1], Rmax=50, type='fast'):
    """CI.calculate_CI (CI.py:107-145): float64 cluster-value array, ``radius * min(vox)`` on
    defect voxels (``type='fast'``; ``'slow'`` leaves table units, as the reference does)."""
    cv, _ = _run(defectArray, vox, Rmax)
    if type == 'slow':
        return cv
    if type == 'fast':
        return cv * np.min(vox)
    raise UnboundLocalError("cannot access local variable 'CI' where it is not associated with a value")
