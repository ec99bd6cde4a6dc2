"""Cluster-mode N4: device time of the resident N4 stage against the size of the chain team
(VB_N4_CL_CHAIN_WARPS; the phase team of a CTA has 1 / K of the one-CTA work, the chain is done by every
CTA in full) and against the schedule without overlap (VB_N4_NO_OVERLAP=1), with bit parity against one
CTA per subject.  One JSON line per (shape, batch)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vent_analysis_b200 import _lib, cohort
from vent_analysis_b200.phantom import phantom

CASES = [((128, 128, 14), [2.0, 2.0, 11.5], 1, (8, 4, 2)), ((128, 128, 14), [2.0, 2.0, 11.5], 16, (8, 4)),
         ((208, 256, 18), [1.5, 1.5, 10.0], 1, (8, 4, 2)), ((256, 256, 64), [1.5, 1.5, 3.0], 1, (8, 4, 2))]
CW = tuple(int(x) for x in os.environ.get("CW", "8,12,16,20,24").split(","))
ctx = _lib.context(0)


def timed(dc, reps=2):
    best = 1e30
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        dc.step()
        e1.record()
        torch.cuda.synchronize()
        if r > 0:
            best = min(best, e0.elapsed_time(e1))
    st = dc.stats()
    return round(best, 3), (dc.n4.cpu().numpy().tobytes(), st["n4_iters"].tobytes(), st["n4_last_conv"].tobytes())


for shape, vox, nb, ks in CASES:
    subs = [phantom(shape, vox, 100 + i % 8) for i in range(min(nb, 8))]
    subs = [subs[i % len(subs)] for i in range(nb)]
    imgs, masks = zip(*subs)
    dc = cohort.DeviceCohort(imgs, masks, [vox] * nb, flags=_lib.DO_N4, want_logbias=True)
    ctx.set_n4_cluster(1)
    t1, ref = timed(dc)
    rec = {"shape": list(shape), "batch": nb, "ms_one_cta": t1, "ms": {}, "identical": True}
    for k in ks:
        ctx.set_n4_cluster(k)
        row = {}
        for cw in CW:
            os.environ["VB_N4_CL_CHAIN_WARPS"] = str(cw)
            t, got = timed(dc)
            row[str(cw)] = t
            rec["identical"] = rec["identical"] and got == ref
        del os.environ["VB_N4_CL_CHAIN_WARPS"]
        os.environ["VB_N4_NO_OVERLAP"] = "1"
        t, got = timed(dc)
        row["serial"] = t
        rec["identical"] = rec["identical"] and got == ref
        del os.environ["VB_N4_NO_OVERLAP"]
        rec["ms"][str(k)] = row
    ctx.set_n4_cluster(0)
    print(json.dumps(rec), flush=True)
    del dc
    torch.cuda.empty_cache()
