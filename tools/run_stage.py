"""Runs selected stages of the path on a resident C2 batch a few times (for ncu captures and
stage timing).   python tools/run_stage.py STAGES [B] [reps]     STAGES e.g. vdp, ci, snr, n4, all"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import torch
from vent_analysis_b200 import _lib, cohort

stages = sys.argv[1] if len(sys.argv) > 1 else "vdp"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 296
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
F = {"snr": _lib.DO_SNR, "n4": _lib.DO_N4, "vdp": _lib.DO_VDP, "ci": _lib.DO_CI, "all": cohort.ALL_STAGES}
flags = 0
for s in stages.split("+"):
    flags |= F[s]
cache = f"/tmp/vb_batch_C2_{B}.npz"
if os.path.exists(cache):
    z = np.load(cache)
    u16, masks = list(z["u16"]), list(z["masks"])
else:
    u16, masks = bench.make_batch("C2", range(B), 16)
    np.savez(cache, u16=np.stack(u16), masks=np.stack(masks))
voxs = [bench.subject_geometry("C2", s)[1] for s in range(B)]
dc = cohort.DeviceCohort([u.astype(np.float32) for u in u16], masks, voxs, flags=cohort.ALL_STAGES)
dc.step(); torch.cuda.synchronize()            # n4 / defect arrays exist for the later stages
dc.flags = flags
ctx = dc.ctx
ctx.set_profiling(True)
for _ in range(reps):
    dc.step()
torch.cuda.synchronize()
ms, n = ctx.profile_fetch()
print(os.environ.get("VENTB200_LIB", "default"), {k: round(v / n, 4) for k, v in ms.items()}, "ms per step of", B, "subjects",
      "iters", int(dc.stats()["n4_iters"].sum()))
if os.environ.get("VB_DUMP_INFO"):
    import json
    st = dc.stats()
    it = st["n4_iters"].sum(axis=1).astype(np.int64)
    json.dump({"subjects_per_launch": B, "subject_iterations": int(it.sum()),
               "mask_voxel_iterations": int((it * st["n4_n_mask"].astype(np.int64)).sum())}, open(os.environ["VB_DUMP_INFO"], "w"))
