"""Timing of single large subjects (log image does not fit shared memory -> global/L2 path)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vent_analysis_b200 import _lib, cohort
from vent_analysis_b200.phantom import phantom
ctx = _lib.context(0)
for shape, vox, nb in [((208, 256, 18), [1.5, 1.5, 10.0], 1), ((208, 256, 18), [1.5, 1.5, 10.0], 16), ((256, 256, 64), [1.5, 1.5, 3.0], 1), ((128, 128, 14), [2.0, 2.0, 11.5], 1)]:
    subs = [phantom(shape, vox, 100 + i) for i in range(nb)]
    imgs, masks = zip(*subs)
    for rep in range(2):
        t = time.perf_counter()
        m, out, st = cohort.run_cohort_host(imgs, masks, [vox] * nb, ctx=ctx, flags=_lib.DO_SNR | _lib.DO_N4 | _lib.DO_VDP)
        dt = time.perf_counter() - t
    print(shape, "batch", nb, "mask voxels", int((masks[0] == 1).sum()), "iters", st["n4_iters"][0][:4].tolist(), f"wall {dt*1e3:.1f} ms", "VDP", round(float(m[0, 0]), 3), "CTAs per subject", ctx.n4_cluster_size())
