"""Oracle (TEST INFRASTRUCTURE) for the sphere-growing lookup table --
restates the regeneration branch of ``CI.getSpherePix`` (CI.py:48-61) shell by
shell, exactly as the reference sweeps it (slow: one pass over the
(2R+1)^3 grid per 0.01 radius step).  PINNED: the SHA-256 of the two tables the
reference ships is committed in ``tests/golden/sphere_tables.json`` and the
product generator (``vent_analysis_b200/sphere_table.py``) must reproduce those
bytes; this literal sweep pins the generator for other spacings at small R.
"""
import numpy as np


def sphere_table_sweep(vox, radius):
    radius = int(radius)                                   # CI.py:49
    vox = np.asarray(vox, dtype=np.float64)
    vox = vox / np.min(vox)                                # :51
    a = range(-radius, radius + 1, 1)
    X, Z, Y = np.meshgrid(a, a, a)                         # :52 ('xy' indexing)
    d2 = (X * vox[0]) ** 2 + (Y * vox[1]) ** 2 + (Z * vox[2]) ** 2
    blocks = [np.zeros((1, 4))]                            # :53 seed row = the centre voxel
    for r in np.arange(0, radius, 0.01):                   # :55
        shell = (d2 <= r ** 2) * (d2 > (r - 0.01) ** 2)    # :56
        x, y, z = X[shell], Y[shell], Z[shell]
        blocks.append(np.column_stack((np.repeat(r, len(x)), x, y, z)))  # :60
    return np.vstack(blocks)
