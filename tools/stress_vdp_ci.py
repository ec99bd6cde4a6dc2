"""Stress check: defect masks, linear-binning codes, borders, VDP / VDP_lb and CI of many random
cohort phantoms, GPU batch vs the oracle GIVEN the GPU's N4 volume (bit-exactness claim)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiprocessing as mp
from vent_analysis_b200.phantom import cohort_member

_N4 = {}


def ref(args):
    i, n4 = args
    from oracle import ci_oracle, vdp_oracle
    from vent_analysis_b200.sphere_table import generate
    img, mask, vox = cohort_member(i)
    r = vdp_oracle.vdp(n4, mask.astype(np.float64), vox)
    ci, shell = ci_oracle.calculate_ci(r["defectArray"], vox, generate(vox, 50))
    return dict(defect=r["defectArray"].astype(np.uint8), lb=r["defectArrayLB"].astype(np.uint8),
                border=r["defectBorder"].astype(np.uint8), vdp=r["VDP"], vdp_lb=r["VDP_lb"],
                ci=vdp_oracle.ci_percentile(ci, r["defectArray"]), shell=shell)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    seeds = list(range(5000, 5000 + n))
    from vent_analysis_b200 import cohort
    subs = [cohort_member(i) for i in seeds]
    imgs, masks, voxs = zip(*subs)
    metrics, out, stats = cohort.run_cohort_host(imgs, masks, voxs)
    offs = np.cumsum([0] + [im.size for im in imgs])
    n4s = [out["n4"][offs[k]:offs[k + 1]].reshape(imgs[k].shape) for k in range(n)]
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        refs = pool.map(ref, list(zip(seeds, n4s)), chunksize=2)
    bad = 0
    for k, r in enumerate(refs):
        shp = imgs[k].shape
        sl = slice(offs[k], offs[k + 1])
        d = out["defect"][sl].reshape(shp)
        ok = (np.array_equal(d, r["defect"]) and np.array_equal(out["defect_lb"][sl].reshape(shp), r["lb"])
              and np.array_equal(out["defect_border"][sl].reshape(shp), r["border"])
              and metrics[k, 0] == r["vdp"] and metrics[k, 1] == r["vdp_lb"] and metrics[k, 2] == r["ci"]
              and np.array_equal(out["ci_shell"][sl].reshape(shp)[d > 0].astype(np.int32), r["shell"][d > 0]))
        if not ok:
            bad += 1
            print("MISMATCH seed", seeds[k])
    print(f"subjects {n}: {bad} with any difference in defect mask / LB codes / border / VDP / VDP_lb / CI map / CI")
