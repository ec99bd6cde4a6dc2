"""Small N4 -> VDP -> CI run for compute-sanitizer (one ragged batch incl. degenerate masks)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vent_analysis_b200 import _lib, cohort
from vent_analysis_b200.phantom import phantom
imgs, masks, voxs = [], [], []
for shape, vox, seed in [((33, 29, 5), [1.5, 1.5, 10.0], 41), ((40, 36, 6), [2.0, 2.0, 11.5], 43), ((36, 40, 1), [1.5, 1.5, 10.0], 42)]:
    i, m = phantom(shape, vox, seed); imgs.append(i); masks.append(m); voxs.append(vox)
ctx = _lib.context(0)
p = ctx.default_n4_params(max_iter=(12, 4, 2))
metrics, out, stats = cohort.run_cohort_host(imgs, masks, voxs, ctx=ctx, params=p)
print("sanitize case ok", metrics[:, :3].round(3).tolist(), stats["n4_iters"][:, :3].tolist())
