"""Timing driver (not part of the product): vb_pack6_batch_dev (Vent_Analysis.build4DdataArray for a
batch) on N resident C2-sized subjects; CUDA events on the launching stream."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vent_analysis_b200 import _lib
from vent_analysis_b200.sphere_table import generate

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1776
shape, vox = (128, 128, 14), [2.0, 2.0, 11.5]
ctx = _lib.context(0)
tbl = generate(vox, 50)
h = ctx.ci_table(tbl, shape, key=(tuple(vox), 50, tbl.shape[0]))
nb = len(ctx.ci_radii(h))
V = n * int(np.prod(shape))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
pro = torch.rand(V, device=dev, generator=g); img = torch.rand(V, device=dev, generator=g); n4 = torch.rand(V, device=dev, generator=g)
mask = (torch.rand(V, device=dev, generator=g) < 0.2).to(torch.uint8)
defect = ((torch.rand(V, device=dev, generator=g) < 0.1) & (mask > 0)).to(torch.uint8)
shell = torch.where(defect > 0, torch.randint(0, nb, (V,), device=dev, generator=g), torch.tensor(0xFFFF, device=dev)).to(torch.uint16)
out = torch.empty(6 * V, device=dev, dtype=torch.float32)
shapes = np.ascontiguousarray([shape] * n, dtype=np.int32)
hs = np.full(n, h, dtype=np.int32)
vm = np.full(n, min(vox), dtype=np.float64)
s = torch.cuda.current_stream().cuda_stream
def call():
    rc = ctx.lib.vb_pack6_batch_dev(ctx.h, n, _lib._ptr(shapes), _lib._ptr(hs), _lib._ptr(vm), pro.data_ptr(), img.data_ptr(),
                                    mask.data_ptr(), n4.data_ptr(), defect.data_ptr(), shell.data_ptr(), out.data_ptr(), C.c_void_p(s))
    ctx.check(rc)
for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    call()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
byt = V * (4 + 4 + 1 + 4 + 1 + 2 + 24)
ref = np.zeros(6, np.float32)
o = out[:6 * 1000].cpu().numpy().reshape(-1, 6)
assert np.array_equal(o[:, 1], img[:1000].cpu().numpy()) and np.array_equal(o[:, 2], mask[:1000].cpu().numpy().astype(np.float32))
print(f"pack6: {n} subjects, {ms:.3f} ms per call (incl. the per-call descriptor upload), {byt/1e9:.2f} GB algorithmic "
      f"(40 B/voxel) -> {byt/ms/1e6:.0f} GB/s = {100*byt/ms/1e6/6548.2:.0f} % of the measured HBM peak")
