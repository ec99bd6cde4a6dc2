"""Stress check of the exact convergence emulation: N4 iteration counts and bias fields of many
random phantoms, GPU batch vs the CPU restatement (oracle), one oracle run per host core."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiprocessing as mp
from vent_analysis_b200.phantom import cohort_member


def ref(i):
    from oracle import n4_oracle
    img, mask, vox = cohort_member(i)
    r = n4_oracle.n4(img, mask)
    return i, r["iters"].tolist(), r["logbias"]


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    seeds = list(range(1000, 1000 + n))
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        refs = pool.map(ref, seeds, chunksize=1)
    from vent_analysis_b200 import _lib
    ctx = _lib.context(0)
    subs = [cohort_member(i) for i in seeds]
    shapes = [s[0].shape for s in subs]
    out = ctx.pipeline_host(shapes, _lib.DO_N4, image=np.concatenate([s[0].ravel() for s in subs]),
                            mask=np.concatenate([s[1].ravel() for s in subs]), want_logbias=True)
    bad, worst, off = 0, 0.0, 0
    for k, (i, it_ref, lb_ref) in enumerate(refs):
        V = subs[k][0].size
        it_gpu = out["stats"]["n4_iters"][k][:4].tolist()
        lb = out["logbias"][off:off + V].reshape(shapes[k]).astype(np.float64)
        rel = float(np.abs(np.expm1(lb - lb_ref)).max())
        worst = max(worst, rel)
        if it_gpu != it_ref:
            bad += 1
            print("MISMATCH seed", i, it_gpu, it_ref, "rel", rel)
        off += V
    print(f"subjects {n}: iteration-count mismatches {bad}, worst bias-field relative error {worst:.2e}")
