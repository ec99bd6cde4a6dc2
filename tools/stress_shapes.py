"""Stress check of the shape / alignment dispatch (not part of the product): many subjects of random
shapes in ONE ragged batch through VDP + CI; defect masks, linear-binning codes, borders and p99
against numpy given the kernel's float32 mean (tests/helpers.py), the CI map against the C oracle.
    python tools/stress_shapes.py [N] [--emu]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from oracle import ci_oracle
from vent_analysis_b200 import _lib
from vent_analysis_b200.sphere_table import generate

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 200
if "--emu" in sys.argv:
    from emu.harness import emu_context
    ctx = emu_context()
else:
    ctx = _lib.context(0)
rng = np.random.default_rng(77)
vox = [1.5, 1.5, 10.0]
tbl = generate(vox, 50)
shapes, vols, handles = [], [], []
for i in range(n):
    kind = rng.integers(0, 4)
    n0 = int(rng.choice([32, 64])) if kind == 0 else int(rng.integers(8, 70))
    n1 = int(rng.integers(8, 40)); n2 = int(rng.integers(1, 16))
    if kind <= 2:
        n1 = (n1 + 3) // 4 * 4                                  # word-eligible rows when the offset is aligned
    shape = (n0, n1, n2)
    x = (1000.0 * (1 + 0.25 * rng.standard_normal(shape))).astype(np.float32)
    m = np.zeros(shape, np.uint8)
    m[n0 // 6:n0 - n0 // 6, n1 // 6:n1 - n1 // 6, :] = 1
    m &= (rng.random(shape) < 0.9).astype(np.uint8)
    # a few blobs of low signal -> defects with clusters
    for _ in range(3):
        c = [int(rng.integers(0, s)) for s in shape]
        r = int(rng.integers(2, 6))
        sl = tuple(slice(max(0, c[k] - r), c[k] + r) for k in range(3))
        x[sl] *= 0.3
    shapes.append(shape); vols.append((x, m))
    handles.append(ctx.ci_table(tbl, shape, key=(tuple(vox), 50, tbl.shape[0])))
out = ctx.pipeline_host(shapes, _lib.DO_VDP | _lib.DO_CI, mask=np.concatenate([v[1].ravel() for v in vols]),
                        n4=np.concatenate([v[0].ravel() for v in vols]), handles=handles)
bad, off, nword = 0, 0, 0
for (x, m), shape, st, h in zip(vols, shapes, out["stats"], handles):
    V = x.size
    vals = np.sort(x[m > 0])
    ok = int(st["n_mask"]) == len(vals) and (len(vals) == 0 or st["p99"] == vals[int(len(vals) * .99)])
    if len(vals):
        ok = ok and st["mean"] == np.float32(vals.astype(np.float64).sum() / len(vals))
        d, lb, bd = helpers._vdp_reference_given_stats(x, m, st["mean"], st["p99"])
        ok = ok and np.array_equal(out["defect"][off:off + V].reshape(shape), d)
        ok = ok and np.array_equal(out["defect_lb"][off:off + V].reshape(shape), lb)
        ok = ok and np.array_equal(out["defect_border"][off:off + V].reshape(shape), bd)
        if d.any() and not st["ci_error"]:
            ci_ref, shell_ref = ci_oracle.calculate_ci(d.astype(np.float64), vox, tbl)
            sel = d > 0
            ok = ok and np.array_equal(out["ci_shell"][off:off + V].reshape(shape)[sel].astype(np.int32), shell_ref[sel])
    nword += int((shape[1] * shape[2]) % 4 == 0 and off % 4 == 0)
    if not ok:
        bad += 1
        print("MISMATCH", shape, "offset", off)
    off += V
print(f"subjects {n} (random shapes, {nword} on the word kernels): {bad} with any difference")
