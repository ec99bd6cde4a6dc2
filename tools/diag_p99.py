"""Diagnostic (not part of the product): p99 / mean of vdp_stats_kernel against numpy on the
adversarial volumes of tests/helpers.py."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from vent_analysis_b200 import _lib

ctx = _lib.context(0)
shapes = [(32, 20, 6), (35, 8, 3), (64, 12, 7), (16, 16, 14), (24, 8, 2), (20, 12, 1), (33, 9, 5), (21, 10, 6)]
odd = [(), (), (), (), (np.inf,), (np.nan, -np.inf), (), (np.inf, np.nan)]
vols = [helpers._near_threshold_volume(s, 100 + i, o) for i, (s, o) in enumerate(zip(shapes, odd))]
for rep in range(2):
    out = ctx.pipeline_host(shapes, _lib.DO_VDP, mask=np.concatenate([v[1].ravel() for v in vols]),
                            n4=np.concatenate([v[0].ravel() for v in vols]))
    for (x, m), shape, st in zip(vols, shapes, out["stats"]):
        with np.errstate(all="ignore"):
            vals = np.sort(x[m > 0])
        k = int(len(vals) * .99)
        g = st["p99"]
        rank = np.searchsorted(vals, g) if g == g else -1
        print(rep, shape, "M", len(vals), st["n_mask"], "k", k, "np", vals[k], "gpu", g, "rank of gpu value", rank,
              "mean", st["mean"], flush=True)
for i, ((x, m), shape) in enumerate(zip(vols, shapes)):
    o = ctx.pipeline_host([shape], _lib.DO_VDP, mask=m, n4=x)
    with np.errstate(all="ignore"):
        vals = np.sort(x[m > 0])
    print("alone", shape, vals[int(len(vals) * .99)], o["stats"]["p99"][0])
