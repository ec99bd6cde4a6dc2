"""Diagnostic (not part of the product): N4 convergence traces GPU vs oracle."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import n4_oracle
from vent_analysis_b200 import _lib
from vent_analysis_b200.phantom import CONFIGS, phantom

libpath = sys.argv[1] if len(sys.argv) > 1 else _lib.LIB_PATH
cfgs = sys.argv[2:] or ["C1"]
ctx = _lib.Context(0, lib=_lib.bind(libpath))
for name in cfgs:
    shape, vox, seed = CONFIGS[name]
    img, mask = phantom(shape, vox, seed)
    ctx.set_debug(True)
    out = ctx.pipeline_host([shape], _lib.DO_N4, image=img, mask=mask, want_logbias=True)
    dbg = ctx.n4_debug(0, 200, 11 ** 3)
    ctx.set_debug(False)
    ref = n4_oracle.n4(img, mask)
    it_g = [int(x) for x in out["stats"]["n4_iters"][0][:4]]
    it_r = [int(x) for x in ref["iters"]]
    print(name, libpath.split('/')[-1], "iters gpu", it_g, "ref", it_r)
    print(" phase us:", {k: round(v, 1) for k, v in dbg["phase_us"].items()}, "per-iteration:",
          {k: round(v / max(sum(it_g), 1), 2) for k, v in dbg["phase_us"].items()})
    g = dbg["conv"]
    pos = 0
    for lvl in range(4):
        r = ref["conv"][lvl][:it_r[lvl]]
        gg = g[pos:pos + it_g[lvl]]
        pos += it_g[lvl]
        n = min(len(r), len(gg))
        print(" level", lvl)
        for i in range(n):
            print(f"   {i:3d} gpu {gg[i]:.7f} ref {r[i]:.7f} rel {abs(gg[i]-r[i])/r[i]:.2e}")
        print("   tail gpu", gg[n:], "ref", r[n:])
    lb = out["logbias"].reshape(shape).astype(np.float64)
    rel = np.abs(np.expm1(lb - ref["logbias"]))
    print(" max rel bias err", rel.max(), "in-mask", rel[mask > 0].max())
