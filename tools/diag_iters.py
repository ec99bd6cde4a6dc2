import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import n4_oracle
from vent_analysis_b200 import _lib
from vent_analysis_b200.phantom import phantom
ctx = _lib.context(0)
for shape, vox, seed in [((64, 64, 8), [1.5, 1.5, 10.0], 11), ((40, 36, 6), [1.5, 1.5, 10.0], 5), ((128, 128, 14), [1.5, 1.5, 10.0], 0)]:
    img, mask = phantom(shape, vox, seed)
    ref = n4_oracle.n4(img, mask)
    for rep in range(4):
        out = ctx.pipeline_host([shape], _lib.DO_N4, image=img, mask=mask, want_logbias=True)
        st = out["stats"][0]
        rel = np.abs(np.expm1(out["logbias"].reshape(shape).astype(np.float64) - ref["logbias"])).max()
        print(shape, "rep", rep, "gpu", list(st["n4_iters"][:4]), "ref", list(ref["iters"]), "last conv", st["n4_last_conv"], "ref conv tail", [c[c >= 0][-1] for c in ref["conv"]], "rel", rel)
