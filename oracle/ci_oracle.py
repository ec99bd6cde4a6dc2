"""Oracle (TEST INFRASTRUCTURE) for the cluster index, CI.py:87-145.

Two restatements:

* :func:`calculate_ci_literal` -- numpy, follows the reference operation by
  operation (``px2vec`` on unchecked coordinates, ``np.intersect1d`` on the growing
  prefix at every shell boundary).  As slow as the reference (~10 ms / defect voxel);
  used for small cases and, on a bounded sample, as the timed CPU baseline.
* :func:`calculate_ci` -- the integer-exact flat-index C restatement in
  ``ci_oracle.c`` (seconds for any config).

PINNED against outputs of the reference's own ``CI.calculate_CI``
(``tests/golden/ci_*.npz``).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _build


def radii_indices(table):
    """``getRadiiIndices`` (CI.py:79-85): row indices at which the radius increases."""
    return np.where(np.diff(table[:, 0]) > 0)[0] + 1


def px2vec(i, j, k, shape):
    """CI.py:65-68."""
    return i + (j - 1) * shape[0] + (k - 1) * shape[0] * shape[1]


def calculate_cv_literal(shape, voxel, def_vec, table, bounds=None):
    """``calculate_CV`` (CI.py:87-105) for one defect voxel -> radius (table units).
    Raises ValueError when no shell boundary satisfies the stop rule (CI.py:101-103)."""
    if bounds is None:
        bounds = radii_indices(table)
    sphere_vec = px2vec(table[:, 1] + voxel[0], table[:, 2] + voxel[1], table[:, 3] + voxel[2], shape)
    for ii in bounds:
        if len(np.intersect1d(sphere_vec[:ii], def_vec)) / ii < 0.5:
            return table[ii - 1, 0]
    raise ValueError("max radius reached")


def defect_list(defect):
    """``multi_which`` (CI.py:10-30): (row, col, slice) of non-zero voxels, C order."""
    return np.argwhere(np.asarray(defect) != 0)


def calculate_ci_literal(defect, vox, table, sample=None):
    """``calculate_CI(type='fast')`` (CI.py:107-145) -> float64 CI array [mm].
    ``sample``: optional index array into the defect list -- only those voxels are
    evaluated (used for the bounded CPU-baseline timing)."""
    defect = np.asarray(defect)
    dl = defect_list(defect)
    def_vec = px2vec(dl[:, 0], dl[:, 1], dl[:, 2], defect.shape)
    bounds = radii_indices(table)
    out = np.zeros(defect.shape, dtype=np.float64)
    which = range(len(dl)) if sample is None else sample
    for n in which:
        v = dl[n]
        out[v[0], v[1], v[2]] = calculate_cv_literal(defect.shape, v, def_vec, table, bounds) * np.min(vox)
    return out


def calculate_ci(defect, vox, table):
    """Fast exact oracle.  Returns ``(CIarray f64 [mm], shell int32)``; raises
    ValueError like the reference when a voxel exhausts the table."""
    lib = _build.load()
    d = np.ascontiguousarray(np.asarray(defect) != 0, dtype=np.uint8)
    t = np.ascontiguousarray(table, dtype=np.float64)
    shell = np.empty(d.shape, dtype=np.int32)
    cv = np.empty(d.shape, dtype=np.float64)
    fn = lib.ci_oracle
    fn.restype = ctypes.c_int
    rc = fn(d.ctypes.data_as(ctypes.c_void_p), *(ctypes.c_int(int(n)) for n in d.shape),
            t.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(t.shape[0]),
            shell.ctypes.data_as(ctypes.c_void_p), cv.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("max radius reached")
    return cv * np.min(vox), shell
