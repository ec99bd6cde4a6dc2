"""N4 with a thread-block cluster of K CTAs per subject (vb_set_n4_cluster): device time of the resident
N4 stage for small batches at K = 1 / 2 / 4 / 8 (CUDA events on the launching stream, best of `reps`)
and bit-parity of every K against K = 1 (corrected volume, log bias, iteration counts, last
convergence value).  Writes one JSON line per (shape, batch)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vent_analysis_b200 import _lib, cohort
from vent_analysis_b200.phantom import phantom

CASES = [((128, 128, 14), [2.0, 2.0, 11.5], (1, 4, 16, 18, 32, 37, 64, 74)),
         ((208, 256, 18), [1.5, 1.5, 10.0], (1, 4, 16)),
         ((256, 256, 64), [1.5, 1.5, 3.0], (1, 2))]
if "--quick" in sys.argv:        # a first sanity pass: one small and one C2 subject
    CASES = [((48, 40, 6), [1.5, 1.5, 10.0], (1, 3)), ((128, 128, 14), [2.0, 2.0, 11.5], (1,))]
if "--small" in sys.argv:        # where the break-even volume lies (the automatic choice's size threshold)
    CASES = [((64, 64, 10), [3.0, 3.0, 12.0], (1, 8)), ((96, 96, 12), [2.5, 2.5, 11.5], (1, 8)), ((112, 112, 14), [2.0, 2.0, 11.5], (1, 8))]
reps = int(os.environ.get("REPS", "3"))
KS = (1, 2, 4, 8)
if "--ncu" in sys.argv:          # one launch of the cluster kernel on a C4 subject (for a capture under ncu)
    CASES, KS, reps = [((256, 256, 64), [1.5, 1.5, 3.0], (1,))], (1, 8), 0
if "--auto" in sys.argv:         # what vb_set_n4_cluster(0) picks, and its time against one CTA per subject
    KS = (1, 0)
ctx = _lib.context(0)
cache = {}
for shape, vox, batches in CASES:
    for nb in batches:
        subs = []
        for i in range(nb):
            key = (shape, i % 8)
            if key not in cache:
                cache[key] = phantom(shape, vox, 100 + i % 8)
            subs.append(cache[key])
        imgs, masks = zip(*subs)
        dc = cohort.DeviceCohort(imgs, masks, [vox] * nb, flags=_lib.DO_N4, want_logbias=True)
        rec = {"shape": list(shape), "batch": nb, "mask_voxels": int((masks[0] == 1).sum()), "ms": {}, "k_used": {}, "identical": {}}
        ref = None
        for k in KS:
            ctx.set_n4_cluster(k)
            best = 1e30
            for r in range(reps + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                dc.step()
                e1.record()
                torch.cuda.synchronize()
                if r > 0 or reps == 0:
                    best = min(best, e0.elapsed_time(e1))
            used = ctx.n4_cluster_size()
            st = dc.stats()
            got = (dc.n4.cpu().numpy().tobytes(), dc.logbias.cpu().numpy().tobytes(), st["n4_iters"].tobytes(),
                   st["n4_last_conv"].tobytes(), st["n4_n_mask"].tobytes())
            if k == 1:
                ref = got
                rec["iters"] = st["n4_iters"][0][:4].tolist()
            rec["ms"][str(k)] = round(best, 3)
            rec["k_used"][str(k)] = used
            rec["identical"][str(k)] = bool(got == ref)
        ctx.set_n4_cluster(0)
        print(json.dumps(rec), flush=True)
        del dc
        torch.cuda.empty_cache()
