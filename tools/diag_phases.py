"""Diagnostic (not part of the product): per-phase times of the N4 kernel for ONE C2 subject, level
by level (runs with 1, 2, 3, 4 levels; the debug capture runs the phases one after the other at
full width and always runs the convergence chain), and the wall time of the same subject without
capture.   python tools/diag_phases.py [lib.so] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vent_analysis_b200 import _lib
from vent_analysis_b200.phantom import CONFIGS, phantom

libpath = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] != "-" else _lib.LIB_PATH
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = _lib.Context(0, lib=_lib.bind(libpath))
shape, vox, _ = CONFIGS["C2"]
img, mask = phantom(shape, vox, seed)
img = np.rint(np.clip(img, 0, 65535)).astype(np.float32)
prev, prev_it, prev_wall = None, 0, 0.0
for nl in (1, 2, 3, 4):
    p = ctx.default_n4_params(max_iter=[50] * nl)
    ctx.set_debug(True)
    out = ctx.pipeline_host([shape], _lib.DO_N4, image=img, mask=mask, params=p)
    dbg = ctx.n4_debug(0, 200, 11 ** 3)
    ctx.set_debug(False)
    its = [int(x) for x in out["stats"]["n4_iters"][0][:nl]]
    for _ in range(3):
        ctx.pipeline_host([shape], _lib.DO_N4, image=img, mask=mask, params=p)
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.pipeline_host([shape], _lib.DO_N4, image=img, mask=mask, params=p)
    wall = (time.perf_counter() - t0) / 5 * 1e6
    ph = dbg["phase_us"]
    d = {k: ph[k] - (prev[k] if prev else 0.0) for k in ph}
    n = its[-1]
    print(f"{os.path.basename(libpath)} levels {nl} iters {its}: level {nl} alone ({n} it) per iteration us: "
          + " ".join(f"{k} {d[k] / max(n, 1):.1f}" for k in ("hist", "sharpen", "fit", "update", "chain"))
          + f" | setup {ph['setup']:.0f} finalize {ph['finalize']:.0f} | wall (no capture) {wall:.0f} us, level adds {(wall - prev_wall) / max(n, 1):.1f} us/it")
    prev, prev_wall = ph, wall
