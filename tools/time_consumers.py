"""Times the kernels around the path on the resident bench batch: overlay RGB
(exportDICOM), build4DdataArray packing, k-space reconstruction.  GPU box only.
    python tools/time_consumers.py [B]"""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vent_analysis_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1776
shape = (128, 128, 14)
V1 = int(np.prod(shape)); V = B * V1
ctx = _lib.context(0)
shapes = np.ascontiguousarray([shape] * B, dtype=np.int32)
dev = torch.device("cuda", 0)
n4 = torch.rand(V, device=dev) * 1500
defect = (torch.rand(V, device=dev) < 0.03).to(torch.uint8)
rgb = torch.empty(3 * V, dtype=torch.uint8, device=dev)
ksp = torch.randn(2 * V, device=dev)
img = torch.empty(2 * V, device=dev)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
sp = C.c_void_p(shapes.ctypes.data)

def timeit(name, fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:46s} {ms:8.3f} ms per {B} subjects  {nbytes / ms / 1e6:8.0f} GB/s algorithmic = {100 * nbytes / ms / 1e6 / peak:5.1f} % of measured HBM peak", flush=True)

p = lambda t: C.c_void_p(t.data_ptr())
timeit("overlay RGB (minmax 4V + compose 5V in, 3V out)",
       lambda: ctx.check(ctx.lib.vb_overlay_rgb_batch_dev(ctx.h, B, sp, p(n4), p(defect), 0, 0, p(rgb), stream)), 12 * V)
timeit("k-space recon (2 passes x (8V in + 8V out))",
       lambda: ctx.check(ctx.lib.vb_recon_batch_dev(ctx.h, B, sp, p(ksp), p(img), stream)), 32 * V)
