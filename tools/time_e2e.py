"""Times the host-buffer entry points on the bench batch in their variants (which inputs, which
outputs), to see what each byte / kernel of the host pipeline costs.  GPU box only.
    python tools/time_e2e.py [B] [steps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vent_analysis_b200 import _lib, cohort

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1776
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
u16, masks = bench.make_batch("C2", range(B), 16)
ctx = _lib.context(0)
shapes = np.ascontiguousarray([u.shape for u in u16], dtype=np.int32)
voxs = [bench.subject_geometry("C2", s)[1] for s in range(B)]
handles = cohort.table_handles(ctx, shapes, voxs)
V = int(sum(u.size for u in u16))
h_u16 = ctx.pinned_empty(V, np.uint16); h_u16[:] = np.concatenate([u.ravel() for u in u16])
h_f32 = ctx.pinned_empty(V, np.float32); h_f32[:] = h_u16
bits = _lib.pack_mask_bits(masks)
h_bits = ctx.pinned_empty(bits.size, np.uint32); h_bits[:] = bits
h_m8 = ctx.pinned_empty(V, np.uint8); h_m8[:] = np.concatenate([m.ravel() for m in masks])
bufs = {"n4": ctx.pinned_empty(V, np.float32), "codes": ctx.pinned_empty(V, np.uint8), "stats": ctx.pinned_empty(B, _lib.STATS_DTYPE),
        "ci_index": ctx.pinned_empty(V // 8, np.uint32), "ci_value": ctx.pinned_empty(V // 8, np.uint16),
        "ci_offset": ctx.pinned_empty(B + 1, np.int64)}
dense = {k: ctx.pinned_empty(V, dt) for k, dt in (("n4", np.float32), ("defect", np.uint8), ("defect_lb", np.uint8),
                                                   ("defect_border", np.uint8), ("ci_shell", np.uint16))}
F = cohort.ALL_STAGES

def timeit(name, fn):
    fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    print(f"{name:60s} {1e3 * dt:8.2f} ms/step  {B / dt:9.0f} subjects/s", flush=True)

timeit("classic f32+u8 in, stats only", lambda: ctx.pipeline_host(shapes, F, image=h_f32, mask=h_m8, handles=handles, stats_only=True))
timeit("classic f32+u8 in, dense arrays out (round-1 e2e)", lambda: ctx.pipeline_host(shapes, F, image=h_f32, mask=h_m8, handles=handles, buffers=dense))
kw = dict(handles=handles, buffers=bufs)
timeit("packed f32+u8 in, stats only", lambda: ctx.pipeline_host_packed(shapes, F, image_f32=h_f32, mask_u8=h_m8, want_n4=False, want_codes=False, want_ci_map=False, **kw))
timeit("packed u16+bits in, stats only", lambda: ctx.pipeline_host_packed(shapes, F, image_u16=h_u16, mask_bits=h_bits, want_n4=False, want_codes=False, want_ci_map=False, **kw))
timeit("packed u16+bits in, codes + sparse CI out (no n4)", lambda: ctx.pipeline_host_packed(shapes, F, image_u16=h_u16, mask_bits=h_bits, want_n4=False, **kw))
timeit("packed u16+bits in, n4 + codes + sparse CI out (e2e)", lambda: ctx.pipeline_host_packed(shapes, F, image_u16=h_u16, mask_bits=h_bits, **kw))
