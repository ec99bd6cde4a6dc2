"""Host <-> device copy bandwidth of the box with N ranks copying at the same time (one process per
GPU under torchrun): H2D alone, D2H alone and both directions together, page-locked buffers allocated
after binding the process to the CPUs next to its GPU (cohort.bind_near_gpu) or unbound.
    python -m torch.distributed.run --nproc-per-node N tools/probe_hostbw.py [--mb 512] [--nobind]
Rank 0 prints one JSON line with the per-rank and aggregate GB/s: the ceiling the e2e numbers of
bench.py run into."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from vent_analysis_b200 import cohort

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=512)
ap.add_argument("--nobind", action="store_true")
ap.add_argument("--reps", type=int, default=6)
a = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
node = None if a.nobind else cohort.bind_near_gpu(local)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = a.mb << 20
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out.fill_(2)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.ones(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def run(up, down):
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.reps):
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    barrier()
    dt = time.perf_counter() - t0
    return a.reps * n * (int(up) + int(down)) / dt / 1e9

res = {}
for name, up, down in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    run(up, down)
    g = run(up, down)
    t = torch.tensor([g], dtype=torch.float64, device="cuda")
    if world > 1:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        per = [float(p) for p in parts]
    else:
        per = [g]
    res[name] = {"aggregate_gbs": round(sum(per), 1), "per_rank_gbs": [round(x, 1) for x in per]}
nodes = torch.tensor([-1 if node is None else node], dtype=torch.int64, device="cuda")
if world > 1:
    parts = [torch.empty_like(nodes) for _ in range(world)]
    dist.all_gather(parts, nodes)
    nodes = [int(p) for p in parts]
else:
    nodes = [int(nodes)]
if rank == 0:
    print(json.dumps({"ranks": world, "mb_per_copy": a.mb, "bound_to_numa_node": nodes, "cpus_rank0": len(os.sched_getaffinity(0)), **res}), flush=True)
if world > 1:
    dist.destroy_process_group()
