#!/usr/bin/env python
"""Benchmark of the N4 -> VDP -> CI voxel hot path (BASELINE.json metric: subject volumes/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = one pass of the whole path (SNR + N4 + VDP + linear binning + borders + CI +
percentile) over one batch of B synthetic subjects per GPU (BASELINE.json configs[1]: seeded
xenon phantoms, 128x128x14 @ 2.0x2.0x11.5 mm; every subject is a distinct phantom).  Subjects
are independent, so ranks shard them with no data-path collective (weak scaling: B per GPU);
the per-subject metric blocks are all-gathered at the end of the run.

Prints ONE JSON line (rank 0):
  value     whole-job subjects/s with the batch resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the host-buffer C-ABI call (pinned host arrays in, every output
            array of the reference's class surface copied back) -- copies inside the timed region
  roofline  the dominant kernel (n4_subject_kernel): algorithmic bytes / its CUDA-event duration
            measured inside the timed region, against MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference CPU path timed on this box's host cores (N=1 only)

``--impl reference`` times the reference's CPU algorithm instead (the reference is pure Python +
SimpleITK and cannot travel to the GPU box, so this is the oracle port: oracle/n4_oracle.c for
SimpleITK's N4, oracle/vdp_oracle.py and the literal np.intersect1d sphere search of
oracle/ci_oracle.py, one subject per host core).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "subject volumes/sec (N4+VDP+CI)"
UNIT = "subjects/s"
SHAPE, VOX = (128, 128, 14), [2.0, 2.0, 11.5]          # BASELINE.json configs[1]
WORKLOAD = "C2 xenon phantoms 128x128x14 @2.0x2.0x11.5 mm, N4(4 levels, 50 it max)+SNR+VDP+LB+borders+CI(Rmax 50)"


# ------------------------------------------------------------------------------------------ inputs
MIXED = False      # --workload C5: the cohort config, C1 / C2 geometry and spacing alternating


def _subject_vox(seed):
    from vent_analysis_b200.phantom import CONFIGS
    return list(CONFIGS["C1" if int(seed) % 2 == 0 else "C2"][1]) if MIXED else list(VOX)


def _make_subject(seed):
    from vent_analysis_b200.phantom import phantom
    return phantom(SHAPE, _subject_vox(seed), int(seed))


def make_batch(seeds, procs):
    import multiprocessing as mp
    if procs > 1 and len(seeds) > 4:
        with mp.get_context("fork").Pool(procs) as pool:
            subs = pool.map(_make_subject, list(seeds), chunksize=max(1, len(seeds) // (4 * procs)))
    else:
        subs = [_make_subject(s) for s in seeds]
    imgs, masks = zip(*subs)
    return list(imgs), list(masks)


# ---------------------------------------------------------------------------- reference CPU path
def _cpu_subject(args):
    """The reference's algorithm for ONE subject on one core (oracle port).  CI runs the
    reference's literal per-voxel np.intersect1d search on `ci_sample` defect voxels (all of
    them when ci_sample <= 0) and its time is scaled to the subject's defect count."""
    seed, ci_sample = args
    from oracle import ci_oracle, n4_oracle, vdp_oracle
    from vent_analysis_b200.sphere_table import generate
    img, mask = _make_subject(seed)
    vox = _subject_vox(seed)
    table = generate(vox, 50)
    t0 = time.perf_counter()
    snr = vdp_oracle.snr(img, mask)
    n4 = n4_oracle.n4(img, mask)
    r = vdp_oracle.vdp(n4["out"], mask.astype(np.float64), vox)
    vdp_oracle.border(mask)
    t1 = time.perf_counter()
    d = r["defectArray"]
    D = int(d.sum())
    if ci_sample > 0 and ci_sample < D:
        sample = np.random.default_rng(seed).choice(D, ci_sample, replace=False)
    else:
        sample = None
    ci_oracle.calculate_ci_literal(d, vox, table, sample=sample)
    t2 = time.perf_counter()
    n_eval = D if sample is None else len(sample)
    t_ci = (t2 - t1) * (D / max(n_eval, 1))
    return dict(t_n4_vdp=t1 - t0, t_ci_measured=t2 - t1, t_ci_full=t_ci, D=D, n_eval=n_eval, snr=float(snr),
                vdp=float(r["VDP"]))


def cpu_step(pool, seeds, ci_sample):
    """One reference step: len(seeds) subjects, one per worker.  Returns (seconds the step takes
    on these cores with the full CI, measured wall seconds, per-subject records)."""
    t0 = time.perf_counter()
    recs = pool.map(_cpu_subject, [(s, ci_sample) for s in seeds], chunksize=1)
    wall = time.perf_counter() - t0
    # every worker handles one subject; the step ends when the slowest one does
    full = max(r["t_n4_vdp"] + r["t_ci_full"] for r in recs)
    full = max(full, wall) if ci_sample <= 0 else full
    return full, wall, recs


def run_reference(args, rank, world):
    """`--impl reference`: rank 0 only."""
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import _build
    _build.build()
    cores = os.cpu_count() or 1
    est_subject_s = 30.0
    budget_s = 150.0
    total_steps = args.steps
    ci_sample = 0 if total_steps * est_subject_s <= budget_s else max(32, int(1250 * budget_s / (total_steps * est_subject_s)))
    seeds0 = 100000
    with mp.get_context("fork").Pool(cores) as pool:
        for w in range(args.warmup):                          # untimed: page in numpy / the oracle .so
            cpu_step(pool, [seeds0 + w * cores + i for i in range(cores)], 8)
        times, walls, recs_all = [], [], []
        for k in range(args.steps):
            full, wall, recs = cpu_step(pool, [seeds0 + (args.warmup + k) * cores + i for i in range(cores)], ci_sample)
            times.append(full); walls.append(wall); recs_all += recs
    total = float(np.sum(times))
    value = cores * args.steps / total
    sample = (f"{cores} subjects per step (one per core), N4/VDP/SNR/borders in full; CI literal search on "
              + ("every defect voxel" if ci_sample <= 0 else
                 f"{ci_sample} sampled defect voxels per subject, its time scaled to the subject's defect count "
                 f"(mean {np.mean([r['D'] for r in recs_all]):.0f}); measured wall {np.sum(walls):.1f} s"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "subjects_per_step": cores,
                   "note": "reference is pure Python + SimpleITK (absent, no network): oracle port of its CPU algorithm"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg():
    """Bounded CPU sample for the `cpu_baseline` object of our own arm: one step of `cores`
    subjects with the full CI (about 15 s)."""
    import multiprocessing as mp
    from oracle import _build
    _build.build()
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        cpu_step(pool, [200000 + i for i in range(cores)], 4)                 # page-in
        full, wall, recs = cpu_step(pool, [300000 + i for i in range(cores)], 0)
    return {"value": cores / full, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} C2 subjects, one per core, full path incl. literal CI on every defect voxel "
                      f"(mean {np.mean([r['D'] for r in recs]):.0f} per subject); wall {wall:.1f} s; N4 is the C restatement "
                      "of SimpleITK's filter (SimpleITK itself is not installable here)"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi in loop mode; samples are kept with their time stamps so that only those that
    fall inside the timed region are used (all samples taken under load if the region was too
    short to catch one)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = str(gpu_index)
        self.path = f"/tmp/vb_clocks_{os.getpid()}.csv"
        self.proc = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", self.gpu], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        rows = []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in open(self.path):
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(p[2]), float(p[3]), [nm for nm, v in zip(names, p[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        try:
            os.remove(self.path)
        except OSError:
            pass
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        use = inside if inside else rows
        if use:
            out.update(sm_mhz=float(np.median([r[1] for r in use])), sm_max_mhz=float(max(r[2] for r in use)),
                       reasons=sorted({x for r in use for x in r[3]}), samples=len(use),
                       window="timed region" if inside else "warm-up + timed region")
        return out


# -------------------------------------------------------------------------------------------- ours
def run_ours(args, rank, local_rank, world):
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_leg()                        # before CUDA is touched (fork pools)
    B = args.batch
    imgs, masks = make_batch([rank * B + i for i in range(B)], min(os.cpu_count() or 1, 16) // max(1, min(world, 8)) or 1)

    import torch
    import torch.distributed as dist
    from vent_analysis_b200 import _lib, cohort
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = _lib.context(local_rank)
    voxs = [_subject_vox(rank * B + i) for i in range(B)]
    dc = cohort.DeviceCohort(imgs, masks, voxs, device=local_rank)
    V = dc.V

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: batch resident in HBM --------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        dc.step()
    barrier()
    ctx.set_profiling(True)
    wall0 = time.time()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        dc.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    stage_ms, ncalls = ctx.profile_fetch()
    ctx.set_profiling(False)
    clocks = sampler.stop(wall0, time.time()) if sampler else None
    stats = dc.stats()
    metrics = dc.metrics()

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -------------------
    def pinned(n, dtype):
        t = torch.empty(n, dtype=dtype, pin_memory=True)
        return t, t.numpy()
    keep = []
    t_img, h_img = pinned(V, torch.float32); t_msk, h_msk = pinned(V, torch.uint8)
    h_img[:] = np.concatenate([im.ravel() for im in imgs]); h_msk[:] = np.concatenate([m.ravel() for m in masks])
    bufs = {}
    for key, dt in (("n4", torch.float32), ("defect", torch.uint8), ("defect_lb", torch.uint8),
                    ("defect_border", torch.uint8), ("ci_shell", torch.uint16)):
        t, a = pinned(V, dt); keep.append(t); bufs[key] = a
    shapes = dc.shapes
    handles = dc.handles
    flags = cohort.ALL_STAGES

    def e2e_step():
        out = ctx.pipeline_host(shapes, flags, image=h_img, mask=h_msk, handles=handles, buffers=bufs)
        assert out["rc"] == 0
        return out
    for _ in range(max(1, min(args.warmup, 2))):
        out_h = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_h = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    # the same call returning only the per-subject records (cohort screening): extra information
    def e2e_metrics_step():
        return ctx.pipeline_host(shapes, flags, image=h_img, mask=h_msk, handles=handles, stats_only=True)
    e2e_metrics_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_m = e2e_metrics_step()
    barrier()
    e2e_m_s = time.perf_counter() - t0
    assert np.array_equal(out_m["stats"]["n_defect"], stats["n_defect"]) and np.array_equal(out_m["stats"]["ci_t95"], stats["ci_t95"])
    h2d = V * 5
    d2h = V * (4 + 1 + 1 + 1 + 2) + B * _lib.STATS_DTYPE.itemsize
    # the two paths must agree bit for bit
    assert np.array_equal(out_h["stats"]["n_defect"], stats["n_defect"]) and np.array_equal(out_h["stats"]["ci_t95"], stats["ci_t95"])

    # ---- reduce over ranks --------------------------------------------------------------------
    t = torch.tensor([ms, e2e_s * 1e3, stage_ms["n4"], e2e_m_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allm = cohort.gather_metrics(metrics, B * world, rank, world)      # final metric gather (NCCL)
    else:
        allm = metrics
    ms, e2e_ms, n4_ms, e2e_m_ms = (float(x) for x in t.cpu())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    n_it = stats["n4_iters"].sum(axis=1).astype(np.float64)
    vox_per = float(np.prod(SHAPE))
    n4_bytes = float(((17.0 + 19.0 * n_it) * vox_per).sum())             # SURVEY.md §8d, per launch (= per step)
    n4_launch_ms = n4_ms / max(ncalls, 1)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    achieved = n4_bytes / (n4_launch_ms * 1e-3) / 1e9
    # DRAM traffic of the kernel from the committed ncu --set full capture (per subject; a launch
    # handles B subjects and the traffic is proportional to them)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_n4_dram_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["subjects_per_launch"] * B
        traffic_src = f"profiles/r01_n4_dram_traffic.json (ncu capture of a {tj['subjects_per_launch']}-subject launch, scaled to {B})"
    line = {
        "metric": METRIC, "value": B * world * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "subjects_per_gpu_per_step": B, "global_subjects_per_step": B * world,
                   "parallelism": f"subject-sharded x{world}, no data-path collective, final all_gather of [n,6] metrics",
                   "l2": f"inputs larger than L2: {V * 5 / 1e6:.0f} MB in + {V * 9 / 1e6:.0f} MB out per GPU per step",
                   "mean_n4_iterations": float(n_it.mean()),
                   "mean_defect_voxels": float(stats["ci_n"].mean())},
        "e2e": {"value": B * world * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
                "api": "vb_pipeline_batch_host (cohort.run_cohort_host), pinned host buffers"},
        "e2e_metrics_only": {"value": B * world * args.steps / (e2e_m_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                             "d2h_bytes_per_step": int(B * _lib.STATS_DTYPE.itemsize), "ms_per_step": e2e_m_ms / args.steps,
                             "note": "same call with VB_HOST_STATS_ONLY: only the per-subject records come back"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "n4_subject_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": n4_bytes, "launch_ms": n4_launch_ms,
                     "definition": "(17 + 19*n_it)*V bytes per subject (SURVEY.md 8d), n_it = iterations actually run"},
        "stage_ms_per_step": {k: v / max(ncalls, 1) for k, v in stage_ms.items()},
        "cpu_baseline": cpu_base,
        "results_check": {"mean_VDP": float(np.nanmean(allm[:, 0])), "mean_CI_mm": float(np.nanmean(allm[:, 2])),
                          "subjects_gathered": int(allm.shape[0])},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1776, help="subjects per GPU per step (12 per SM; the cohort config is 1024)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C2", "C5"],
                    help="C2: BASELINE.json configs[1] (default); C5: the cohort config, C1/C2 spacings alternating")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global MIXED, WORKLOAD
    if args.workload == "C5":
        MIXED = True
        WORKLOAD = ("C5 cohort: xenon phantoms 128x128x14, spacings 1.5x1.5x10 / 2.0x2.0x11.5 mm alternating, "
                    "N4(4 levels, 50 it max)+SNR+VDP+LB+borders+CI(Rmax 50)")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
