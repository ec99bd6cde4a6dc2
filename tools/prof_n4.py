"""Profiling driver (not part of the product): one N4 launch over a batch of C2 phantoms."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vent_analysis_b200 import _lib
from vent_analysis_b200.phantom import phantom

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
shape, vox = (128, 128, 14), [2.0, 2.0, 11.5]
subs = [phantom(shape, vox, i) for i in range(min(n, 8))]
imgs = [subs[i % len(subs)][0] for i in range(n)]
masks = [subs[i % len(subs)][1] for i in range(n)]
ctx = _lib.context(0)
img = np.concatenate([a.ravel() for a in imgs]); msk = np.concatenate([m.ravel() for m in masks])
for _ in range(reps):
    out = ctx.pipeline_host([shape] * n, _lib.DO_N4, image=img, mask=msk)
print("iters", out["stats"]["n4_iters"][0][:4], "launches", ctx.launch_count())
