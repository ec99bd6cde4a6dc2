"""Sphere-growing lookup tables for the cluster index (reference:
``CI.getSpherePix``, CI.py:33-63; shipped as ``1.5x1.5x10.0_50.npy`` and
``2.0x2.0x11.5_50.npy``).

A table is an ``N x 4`` float64 array ``[radius, dx, dy, dz]`` sorted by radius:
row 0 is the centre voxel, then one block per 0.01-wide shell (radius in units of
the smallest voxel edge) listing the integer voxel offsets whose scaled distance
``d`` satisfies ``(r-0.01)^2 < d^2 <= r^2``.  The reference builds it with 100*R
full sweeps over a (2R+1)^3 grid (minutes); here every grid point is assigned to
its shell(s) directly with the *same* float64 expressions, so the result is
byte-identical (including the four offsets (+-32,0,0),(0,+-32,0) that fall in two
consecutive shells because ``(r-0.01)**2`` is not exactly the previous ``r**2``).
``tests/test_sphere_table.py`` checks the bytes against the SHA-256 of the
reference's shipped tables.
"""
from __future__ import annotations

import os

import numpy as np

_CACHE: dict = {}


def table_file_name(vox, radius=50):
    """File name the reference looks up in ``os.getcwd()`` (CI.py:43)."""
    return f'{vox[0]}x{vox[1]}x{vox[2]}_{radius}.npy'


def generate(vox, radius=50):
    """Exact table for voxel size ``vox`` (row, col, slice) up to ``radius``."""
    radius = int(radius)
    v = np.asarray(vox, dtype=np.float64)
    v = v / np.min(v)
    a = np.arange(-radius, radius + 1, 1)
    # CI.py:52 -- X varies along axis 1, Z along axis 0, Y along axis 2; a boolean
    # selection therefore lists a shell z-major, then x, then y.
    X, Z, Y = np.meshgrid(a, a, a)
    d2 = ((X * v[0]) ** 2 + (Y * v[1]) ** 2 + (Z * v[2]) ** 2).ravel()
    rs = np.arange(0, radius, 0.01)
    hi2 = rs ** 2
    lo2 = (rs - 0.01) ** 2
    k0 = np.searchsorted(hi2, d2, side='left')
    gid = np.arange(d2.size)
    ks, ids = [], []
    for dk in (-2, -1, 0, 1, 2):          # rounding can move a point by at most one shell
        k = k0 + dk
        ok = (k >= 0) & (k < rs.size)
        kk = np.clip(k, 0, rs.size - 1)
        hit = ok & (d2 <= hi2[kk]) & (d2 > lo2[kk])
        ks.append(kk[hit])
        ids.append(gid[hit])
    ks = np.concatenate(ks)
    ids = np.concatenate(ids)
    o = np.lexsort((ids, ks))
    ks, ids = ks[o], ids[o]
    out = np.empty((ks.size + 1, 4), dtype=np.float64)
    out[0] = 0.0
    out[1:, 0] = rs[ks]
    out[1:, 1] = X.ravel()[ids]
    out[1:, 2] = Y.ravel()[ids]
    out[1:, 3] = Z.ravel()[ids]
    return out


def get_sphere_pix(vox, radius=50, search_cwd=True):
    """Drop-in for ``CI.getSpherePix``: load ``<vox>_<radius>.npy`` from the working
    directory when it exists (CI.py:44-46), else generate (and memoise) it.  Unlike
    the reference nothing is written to the working directory."""
    name = table_file_name(vox, radius)
    if search_cwd:
        path = os.path.join(os.getcwd(), name)
        if os.path.exists(path):
            try:
                return np.load(path)
            except Exception:
                pass
    key = (tuple(float(x) for x in vox), int(radius))
    if key not in _CACHE:
        _CACHE[key] = generate(vox, radius)
    return _CACHE[key]
