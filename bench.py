#!/usr/bin/env python
"""Benchmark of the N4 -> VDP -> CI voxel hot path (BASELINE.json metric: subject volumes/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1|C2|C3|C4|C5] [--batch B]
                    [--scaling weak|strong] [--cohort 1024] [--impl ours|reference]

One "step" = one pass of the whole path (SNR + N4 + VDP + linear binning + borders + CI +
percentile) over one batch of synthetic subjects per GPU.  Every subject is a distinct seeded
phantom whose pixel values are integers, as DICOM pixel data are (the reference reads uint16,
Vent_Analysis.py:169-181, and casts to float32 at :322).  Default workload: BASELINE.json
configs[1] (C2: 128x128x14 @ 2.0x2.0x11.5 mm), 1776 subjects per GPU, weak scaling (subjects are
independent, ranks shard them with no data-path collective; the per-subject metric blocks are
all-gathered at the end of the run).  ``--workload`` selects the other configs (C1, C3: 128^3 with
the regenerated isotropic table, C4: 256x256x64 N4 + linear binning without CI, C5: the mixed-spacing
cohort); ``--scaling strong --cohort 1024`` splits a fixed cohort over the ranks.

Prints ONE JSON line (rank 0):
  value     whole-job subjects/s with the batch resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the host-buffer C-ABI call vb_pipeline_batch_host_packed: pinned
            host arrays in (uint16 image, bit-packed mask), N4HPvent + one code byte per voxel + the
            sparse CI list + per-subject records out -- every copy inside the timed region
  roofline  the dominant kernel (n4_subject_kernel): algorithmic bytes / its CUDA-event duration
            measured inside the timed region, against MEASURED_PEAKS.json; ``onchip`` puts the same
            launch against the SM's issue-slot and shared-memory ceilings (the kernel's real bound)
  cpu_baseline  the oracle port of the reference CPU path timed on this box's host cores (N=1 only)
  cohort_c5 (default run only) the 1024-subject mixed-spacing cohort split over the ranks

``--impl reference`` times the reference's CPU algorithm instead (the reference is pure Python +
SimpleITK and cannot travel to the GPU box, so this is the oracle port: oracle/n4_oracle.c for
SimpleITK's N4, oracle/vdp_oracle.py and the literal np.intersect1d sphere search of
oracle/ci_oracle.py on EVERY defect voxel, one subject per host core, wall-clock timed).
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
STAGES_FULL = "N4(4 levels, 50 it max)+SNR+VDP+LB+borders+CI(Rmax 50)"
# name -> (description, default subjects per GPU per step, runs CI)
WORKLOADS = {
    "C1": ("C1 xenon phantoms 128x128x14 @1.5x1.5x10 mm, " + STAGES_FULL, 1776, True),
    "C2": ("C2 xenon phantoms 128x128x14 @2.0x2.0x11.5 mm, " + STAGES_FULL, 1776, True),
    "C3": ("C3 isotropic xenon phantoms 128x128x128 @3.9 mm, CI table regenerated for that spacing (523k rows), " + STAGES_FULL, 148, True),
    "C4": ("C4 high-res volumes 256x256x64, N4(4 levels, 50 it max)+SNR+VDP+LB+borders, no CI (BASELINE.json configs[3])", 148, False),
    "C5": ("C5 cohort: xenon phantoms 128x128x14, spacings 1.5x1.5x10 / 2.0x2.0x11.5 mm alternating, " + STAGES_FULL, 1776, True),
}


def subject_geometry(workload, seed):
    from vent_analysis_b200.phantom import CONFIGS
    name = workload if workload != "C5" else ("C1" if int(seed) % 2 == 0 else "C2")
    shape, vox, _ = CONFIGS[name]
    return tuple(shape), list(vox)


def _make_subject(args):
    """(uint16 pixel data, uint8 mask): the seeded phantom rounded to integers like DICOM pixels."""
    from vent_analysis_b200.phantom import phantom
    workload, seed = args
    shape, vox = subject_geometry(workload, seed)
    img, mask = phantom(shape, vox, int(seed))
    return np.rint(np.clip(img, 0, 65535)).astype(np.uint16), mask


def make_batch(workload, seeds, procs):
    import multiprocessing as mp
    jobs = [(workload, int(s)) for s in seeds]
    if procs > 1 and len(jobs) > 4:
        with mp.get_context("fork").Pool(procs) as pool:
            subs = pool.map(_make_subject, jobs, chunksize=max(1, len(jobs) // (4 * procs)))
    else:
        subs = [_make_subject(j) for j in jobs]
    u16, masks = zip(*subs)
    return list(u16), list(masks)


# ---------------------------------------------------------------------------- reference CPU path
def _cpu_subject(args):
    """The reference's algorithm for ONE subject on one core (oracle port), wall-clock timed.  CI is
    the reference's literal per-voxel np.intersect1d search on every defect voxel (`ci_sample` > 0
    limits it to that many voxels: warm-up only)."""
    workload, seed, ci_sample = args
    from oracle import ci_oracle, n4_oracle, vdp_oracle
    from vent_analysis_b200.sphere_table import generate
    u16, mask = _make_subject((workload, seed))
    img = u16.astype(np.float32)
    _, vox = subject_geometry(workload, seed)
    table = generate(vox, 50) if WORKLOADS[workload][2] else None
    t0 = time.perf_counter()
    snr = vdp_oracle.snr(img, mask)
    n4 = n4_oracle.n4(img, mask)
    r = vdp_oracle.vdp(n4["out"], mask.astype(np.float64), vox)
    vdp_oracle.border(mask)
    t1 = time.perf_counter()
    d = r["defectArray"]
    D = int(d.sum())
    if table is not None:
        sample = None
        if 0 < ci_sample < D:
            sample = np.random.default_rng(seed).choice(D, ci_sample, replace=False)
        ci_oracle.calculate_ci_literal(d, vox, table, sample=sample)
    t2 = time.perf_counter()
    return dict(t_n4_vdp=t1 - t0, t_ci=t2 - t1, D=D, snr=float(snr), vdp=float(r["VDP"]))


def cpu_step(pool, workload, seeds, ci_sample=0):
    """One reference step: len(seeds) subjects, one per worker.  Returns (wall seconds, records)."""
    t0 = time.perf_counter()
    recs = pool.map(_cpu_subject, [(workload, s, ci_sample) for s in seeds], chunksize=1)
    return time.perf_counter() - t0, recs


def run_reference(args, rank, world):
    """`--impl reference`: rank 0 only.  Every step is `cores` subjects (one per core) through the
    whole path with the literal CI on every defect voxel, timed by the wall clock; the number of
    steps is cut to what fits the time budget (a step takes ~10 s) and the line says how many ran."""
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import _build
    _build.build()
    cores = os.cpu_count() or 1
    budget_s = float(os.environ.get("VB_REF_BUDGET_S", "150"))
    wl = args.workload
    seeds0 = 100000
    walls, recs_all = [], []
    with mp.get_context("fork").Pool(cores) as pool:
        cpu_step(pool, wl, [seeds0 - 1 - i for i in range(cores)], ci_sample=4)      # untimed: page in numpy / the oracle .so
        t_start = time.perf_counter()
        for k in range(max(1, args.steps)):
            if walls and (time.perf_counter() - t_start) + 1.15 * max(walls) > budget_s:
                break
            wall, recs = cpu_step(pool, wl, [seeds0 + k * cores + i for i in range(cores)])
            walls.append(wall); recs_all += recs
    steps = len(walls)
    total = float(np.sum(walls))
    value = cores * steps / total
    sample = (f"{steps} steps of {cores} subjects (one per core), whole path in full incl. the literal CI search on every defect "
              f"voxel (mean {np.mean([r['D'] for r in recs_all]):.0f} per subject); wall clock {total:.1f} s; "
              f"{args.steps} steps were asked for, {steps} fit the {budget_s:.0f} s budget; warm-up = one untimed short step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[wl][0], "subjects_per_step": cores, "steps_requested": args.steps,
                   "note": "reference is pure Python + SimpleITK (absent, no network): oracle port of its CPU algorithm, wall-clock timed"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(workload):
    """Bounded CPU sample for the `cpu_baseline` object of our own arm: one step of `cores`
    subjects with the full CI (about 10-15 s)."""
    import multiprocessing as mp
    from oracle import _build
    _build.build()
    cores = os.cpu_count() or 1
    if workload in ("C3", "C4"):
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "not timed for this workload: one subject takes minutes on a core (the default C2 run carries the CPU baseline)"}
    with mp.get_context("fork").Pool(cores) as pool:
        cpu_step(pool, workload, [200000 + i for i in range(cores)], ci_sample=4)     # page-in
        wall, recs = cpu_step(pool, workload, [300000 + i for i in range(cores)])
    return {"value": cores / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} {workload} subjects, one per core, full path incl. literal CI on every defect voxel "
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
class Leg:
    """One batch on one GPU: resident (`value`) and host-buffer (`e2e`) measurements."""

    def __init__(self, ctx, workload, seeds, local_rank, data):
        from vent_analysis_b200 import _lib, cohort
        self.ctx, self.workload = ctx, workload
        self.flags = cohort.ALL_STAGES if WORKLOADS[workload][2] else (_lib.DO_SNR | _lib.DO_N4 | _lib.DO_VDP)
        self.u16, self.masks = data
        self.voxs = [subject_geometry(workload, s)[1] for s in seeds]
        self.B = len(seeds)
        self.dc = cohort.DeviceCohort([u.astype(np.float32) for u in self.u16], self.masks, self.voxs, device=local_rank,
                                      flags=self.flags)
        self.V = self.dc.V

    def resident(self, steps, warmup, barrier, sampler=None):
        import torch
        ctx, dc = self.ctx, self.dc
        if sampler:
            sampler.start()
        for _ in range(warmup):
            dc.step()
        barrier()
        ctx.set_profiling(True)
        wall0 = time.time()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            dc.step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ctx.launch_count() - l0
        stage_ms, ncalls = ctx.profile_fetch()
        ctx.set_profiling(False)
        clocks = sampler.stop(wall0, time.time()) if sampler else None
        self.stats = dc.stats()
        self.metrics = dc.metrics()
        return dict(ms=ms, launches=launches, stage_ms=stage_ms, ncalls=ncalls, clocks=clocks)

    def e2e(self, steps, warmup, barrier):
        """Through vb_pipeline_batch_host_packed with page-locked host buffers (vb_host_alloc): uint16
        image + mask bits up; N4HPvent, code bytes, sparse CI list and records down."""
        from vent_analysis_b200 import _lib
        ctx, V, B = self.ctx, self.V, self.B
        shapes = self.dc.shapes
        h_img = ctx.pinned_empty(V, np.uint16)
        h_img[:] = np.concatenate([u.ravel() for u in self.u16])
        bits = _lib.pack_mask_bits(self.masks)
        h_bits = ctx.pinned_empty(bits.size, np.uint32)
        h_bits[:] = bits
        want_ci = bool(self.flags & _lib.DO_CI)
        cap = max(1024, V // 8)                         # defect voxels are a few per cent of the volume
        bufs = {"n4": ctx.pinned_empty(V, np.float32), "codes": ctx.pinned_empty(V, np.uint8),
                "stats": ctx.pinned_empty(B, _lib.STATS_DTYPE)}
        if want_ci:
            bufs.update(ci_index=ctx.pinned_empty(cap, np.uint32), ci_value=ctx.pinned_empty(cap, np.uint16),
                        ci_offset=ctx.pinned_empty(B + 1, np.int64))

        def step(arrays=True):
            out = ctx.pipeline_host_packed(shapes, self.flags, image_u16=h_img, mask_bits=h_bits, handles=self.dc.handles,
                                           buffers=bufs, want_n4=arrays, want_codes=arrays, want_ci_map=arrays and want_ci)
            assert out["rc"] == 0
            return out
        for _ in range(max(1, min(warmup, 2))):
            out = step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = step()
        barrier()
        e2e_s = time.perf_counter() - t0
        D_total = int(out["ci_offset"][B]) if want_ci else 0
        # the two paths must agree bit for bit
        st = self.stats
        assert np.array_equal(out["stats"]["n_defect"], st["n_defect"]) and np.array_equal(out["stats"]["ci_t95"], st["ci_t95"])
        assert np.array_equal(out["stats"]["n4_iters"], st["n4_iters"])
        # the same call returning only the per-subject records (cohort screening): extra information
        step(False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out_m = step(False)
        barrier()
        e2e_m_s = time.perf_counter() - t0
        assert np.array_equal(out_m["stats"]["n_defect"], st["n_defect"])
        h2d = 2 * V + 4 * bits.size
        d2h = 4 * V + V + 6 * D_total + B * _lib.STATS_DTYPE.itemsize
        return dict(e2e_s=e2e_s, e2e_m_s=e2e_m_s, h2d=h2d, d2h=d2h, d2h_m=B * _lib.STATS_DTYPE.itemsize)


def _mean(x):
    """nanmean that stays JSON-clean (None when nothing is finite, e.g. CI for the C4 workload)."""
    x = np.asarray(x, dtype=np.float64)
    return float(np.nanmean(x)) if np.isfinite(x).any() else None


def n4_algorithmic_bytes(stats, shapes):
    """(17 + 19 n_it) V bytes per subject, n_it = iterations actually run."""
    n_it = stats["n4_iters"].sum(axis=1).astype(np.float64)
    vox = np.prod(np.asarray(shapes, dtype=np.float64), axis=1)
    return float(((17.0 + 19.0 * n_it) * vox).sum()), n_it


def run_ours(args, rank, local_rank, world):
    from vent_analysis_b200 import cohort as cohort_mod
    cohort_mod.bind_near_gpu(local_rank)                     # host threads and pinned pages next to this rank's GPU
    wl = args.workload
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_leg(wl)                      # before CUDA is touched (fork pools)
    strong = args.scaling == "strong"
    if strong:
        seeds = cohort_mod.shard_indices(args.cohort, rank, world)
    else:
        B = args.batch if args.batch > 0 else WORKLOADS[wl][1]
        seeds = [rank * B + i for i in range(B)]
    procs = max(1, min(os.cpu_count() or 1, 16) // max(1, min(world, 8)))
    # every phantom is made BEFORE CUDA / NCCL come up: the generator forks worker processes, and a
    # fork of a process that already runs NCCL's threads can deadlock in the child
    data = make_batch(wl, seeds, procs)
    with_cohort = wl == "C2" and not strong and args.cohort_leg and args.batch <= 0
    if with_cohort:
        cseeds = cohort_mod.shard_indices(1024, rank, world)
        cdata = make_batch("C5", cseeds, procs)

    import torch
    import torch.distributed as dist
    from vent_analysis_b200 import _lib
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = _lib.context(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    def reduce_sum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t.cpu()]

    leg = Leg(ctx, wl, seeds, local_rank, data)
    B, V = leg.B, leg.V
    r = leg.resident(args.steps, args.warmup, barrier, ClockSampler(local_rank) if rank == 0 else None)
    e = leg.e2e(args.steps, args.warmup, barrier)
    n4_bytes, n_it = n4_algorithmic_bytes(leg.stats, leg.dc.shapes)
    ms, e2e_ms, n4_ms, e2e_m_ms = reduce_max([r["ms"], e["e2e_s"] * 1e3, r["stage_ms"]["n4"], e["e2e_m_s"] * 1e3])
    n_total, = reduce_sum([float(B)])
    n_total = int(round(n_total))
    allm = cohort_mod.gather_metrics(leg.metrics, n_total, rank, world) if world > 1 and not strong else leg.metrics
    if world > 1 and strong:
        allm = cohort_mod.gather_metrics(leg.metrics, args.cohort, rank, world)

    # ---- the cohort config (BASELINE.json configs[4]) beside the default run: 1024 mixed-spacing
    # subjects split over the ranks (strong scaling), same two measurements
    cohort_c5 = None
    if with_cohort:
        del leg.dc
        torch.cuda.empty_cache()
        cl = Leg(ctx, "C5", cseeds, local_rank, cdata)
        cr = cl.resident(args.steps, args.warmup, barrier)
        ce = cl.e2e(args.steps, args.warmup, barrier)
        c_ms, c_e2e_ms = reduce_max([cr["ms"], ce["e2e_s"] * 1e3])
        cm = cohort_mod.gather_metrics(cl.metrics, 1024, rank, world)
        cohort_c5 = {"workload": WORKLOADS["C5"][0], "scaling": "strong", "cohort_subjects": 1024,
                     "subjects_per_gpu": len(cseeds), "value": 1024 * args.steps / (c_ms * 1e-3), "unit": UNIT,
                     "ms_per_step": c_ms / args.steps,
                     "e2e": {"value": 1024 * args.steps / (c_e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": c_e2e_ms / args.steps,
                             "h2d_bytes_per_step": int(ce["h2d"]), "d2h_bytes_per_step": int(ce["d2h"])},
                     "mean_VDP": _mean(cm[:, 0]), "mean_CI_mm": _mean(cm[:, 2]),
                     "subjects_gathered": int(np.isfinite(cm[:, 0]).sum()),
                     "limiter": "148 persistent CTAs pull whole subjects (about 7 ms each): below one wave of subjects per GPU "
                                "the step is one subject's latency, so strong scaling flattens once 1024 / N approaches 148"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    n4_launch_ms = n4_ms / max(r["ncalls"], 1)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    achieved = n4_bytes / (n4_launch_ms * 1e-3) / 1e9
    roof = {"kernel": "n4_subject_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": n4_bytes, "launch_ms": n4_launch_ms,
            "definition": "(17 + 19*n_it)*V bytes per subject (SURVEY.md 8d), n_it = iterations actually run; launch_ms = CUDA events "
                          "from the first N4 launch of a step to the join of its streams (large batches run in two parts whose "
                          "N4 kernels overlap: the small VDP / CI kernels of the last part are inside this time)"}
    # counters of the committed ncu --set full capture of this kernel (same build, C2 batch): DRAM
    # traffic, warp instructions and shared-memory wavefronts per subject-iteration -- put against
    # the time measured live
    cpath = os.path.join(ROOT, "profiles", "n4_counters.json")
    if os.path.exists(cpath) and wl in ("C1", "C2", "C5"):
        cj = json.load(open(cpath))
        scale = float(n_it.sum()) / cj["subject_iterations"]
        roof["traffic"] = (cj["dram_bytes_read"] + cj["dram_bytes_write"]) / cj["subjects_per_launch"] * B
        roof["traffic_unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)"
        roof["traffic_source"] = f"profiles/n4_counters.json ({cj['source']}), scaled to {B} subjects"
        sm_hz = 1e6 * (r["clocks"]["sm_mhz"] if r["clocks"] and r["clocks"]["sm_mhz"] else 1965.0)
        sms = 148
        inst = cj["warp_instructions"] * scale
        wav = cj["smem_wavefronts"] * scale
        t = n4_launch_ms * 1e-3
        roof["onchip"] = {
            "issue": {"achieved": inst / t / 1e9, "peak": sms * 4 * sm_hz / 1e9, "unit": "G warp-instructions/s",
                      "frac": inst / t / (sms * 4 * sm_hz)},
            "shared_memory": {"achieved": 128 * wav / t / 1e9, "peak": sms * 128 * sm_hz / 1e9, "unit": "GB/s",
                              "frac": wav / t / (sms * sm_hz)},
            "warp_instructions_per_mask_voxel_iteration": cj["warp_instructions"] * 32 / cj["mask_voxel_iterations"],
            "note": "instruction and shared-memory wavefront counts from the ncu capture (deterministic for a given batch), "
                    "scaled by the iterations of this run; peak = 148 SMs x 4 issue slots (128 B of shared memory) per clock "
                    "at the SM clock sampled during the timed region",
        }
    line = {
        "metric": METRIC, "value": n_total * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[wl][0], "subjects_per_gpu_per_step": B, "global_subjects_per_step": n_total,
                   "parallelism": f"subject-sharded x{world}, no data-path collective, final all_gather of [n,6] metrics",
                   "l2": f"inputs larger than L2: {V * 5 / 1e6:.0f} MB in + {V * 9 / 1e6:.0f} MB out per GPU per step (resident arrays)",
                   "pixel_data": "integer-valued phantoms (DICOM-like uint16), float32 on the device",
                   "mean_n4_iterations": float(n_it.mean()),
                   "mean_defect_voxels": float(leg.stats["ci_n"].mean())},
        "e2e": {"value": n_total * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(e["h2d"]),
                "d2h_bytes_per_step": int(e["d2h"]), "ms_per_step": e2e_ms / args.steps,
                "api": "vb_pipeline_batch_host_packed, page-locked host buffers (vb_host_alloc): uint16 image + mask bits in; "
                       "N4HPvent f32 + code byte per voxel + sparse CI list + records out"},
        "e2e_metrics_only": {"value": n_total * args.steps / (e2e_m_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(e["h2d"]),
                             "d2h_bytes_per_step": int(e["d2h_m"]), "ms_per_step": e2e_m_ms / args.steps,
                             "note": "same call with no output arrays: only the per-subject records come back"},
        "gpu_launches": int(r["launches"]),
        "clocks": r["clocks"],
        "roofline": roof,
        "stage_ms_per_step": {k: v / max(r["ncalls"], 1) for k, v in r["stage_ms"].items()},
        "cpu_baseline": cpu_base,
        "results_check": {"mean_VDP": _mean(allm[:, 0]), "mean_CI_mm": _mean(allm[:, 2]),
                          "subjects_gathered": int(np.isfinite(allm[:, 0]).sum())},
    }
    if cohort_c5 is not None:
        line["cohort_c5"] = cohort_c5
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="subjects per GPU per step (default: per workload; 1776 = 12 per SM for C1/C2/C5)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS),
                    help="C2: BASELINE.json configs[1] (default); C1, C3, C4, C5: the other configs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: a fixed cohort of --cohort subjects split over the ranks")
    ap.add_argument("--cohort", type=int, default=1024)
    ap.add_argument("--no-cohort-leg", dest="cohort_leg", action="store_false",
                    help="skip the C5 cohort measurement that rides on the default run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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
