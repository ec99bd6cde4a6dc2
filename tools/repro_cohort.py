import sys, os, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vent_analysis_b200 import cohort, _lib
from vent_analysis_b200.phantom import cohort_member
subs = [cohort_member(i) for i in range(12)]
imgs, masks, voxs = zip(*subs)
try:
    metrics, out, stats = cohort.run_cohort_host(imgs, masks, voxs); print("host ok")
    dc = cohort.DeviceCohort(imgs, masks, voxs); dc.step(); dc.torch.cuda.synchronize(); print("dev ok")
    m2 = dc.metrics(); print("equal", np.array_equal(metrics, m2))
    for i in (0, 5, 11):
        one, o1, _ = cohort.run_cohort_host([imgs[i]], [masks[i]], [voxs[i]]); print("single", i, "ok", np.array_equal(one[0], metrics[i]))
    m3, o3, _ = cohort.run_cohort_host(imgs, masks, voxs); print("again ok", np.array_equal(m3, metrics))
except Exception:
    traceback.print_exc()
