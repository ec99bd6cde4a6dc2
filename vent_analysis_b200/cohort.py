"""Batch ("cohort") entry points: many subjects per library call, sharded per subject over
the GPUs of one box.  A single 128x128x14 volume is 0.9 MB -- far too small to
load a B200 -- so throughput comes from ragged batches: every kernel takes the whole batch and
the N4 kernel lets each SM pull subjects from a work counter.

Host arrays go through ``vb_pipeline_batch_host`` (copies inside); :class:`DeviceCohort` keeps
the batch resident in HBM (allocated with torch, which is plumbing only) and calls
``vb_pipeline_batch_dev`` on torch's current stream.

Multi-GPU: subjects are independent (``calculate_VDP`` / ``calculate_CI`` only touch ``self``),
so rank r takes subjects ``r, r+world, ...`` -- no data-path collective; one
``all_gather`` of the 6-float per-subject metric block at the end.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .sphere_table import get_sphere_pix

METRICS = ("VDP", "VDP_lb", "CI", "SNR", "LungVolume", "DefectVolume")
ALL_STAGES = _lib.DO_SNR | _lib.DO_N4 | _lib.DO_VDP | _lib.DO_CI


def bind_near_gpu(device=0):
    """Restrict this process to the CPUs of the NUMA node the GPU hangs off (Linux sysfs), so that
    the page-locked staging buffers it allocates afterwards (first touch) and the threads that feed
    the copies are local to that GPU's PCIe root.  Returns the node, or None when the box has a
    single node / the topology is not exposed (then nothing is changed)."""
    import glob
    import os
    try:
        nodes = [p for p in glob.glob("/sys/devices/system/node/node[0-9]*")]
        if len(nodes) < 2:
            return None
        import subprocess
        # (nvidia-smi, not the CUDA runtime: callers fork worker processes after this)
        idx = str(device)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis:
            ent = [v.strip() for v in vis.split(",") if v.strip()]
            if device < len(ent):
                idx = ent[device]
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", idx],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bdf:
            return None
        path = f"/sys/bus/pci/devices/{bdf[-12:]}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        return None
    return None


def shard_indices(n, rank, world):
    """Subjects of ``rank``: round robin (CI cost varies per subject; interleaving balances it)."""
    return list(range(int(rank), int(n), int(world)))


def table_handles(ctx, shapes, voxs, rmax=50):
    """One CI table handle per subject (memoised per (vox, shape) inside the context)."""
    hs = []
    for shp, vox in zip(shapes, voxs):
        vox = [float(v) for v in vox]
        t = get_sphere_pix(vox, rmax, search_cwd=False)
        hs.append(ctx.ci_table(t, shp, key=(tuple(vox), int(rmax), t.shape[0])))
    return np.asarray(hs, dtype=np.int32)


def metrics_from_stats(ctx, stats, shapes, voxs, handles, n_mask_eq1=None):
    """[n, 6] float64 (VDP, VDP_lb, CI, SNR, LungVolume, DefectVolume), formed exactly as the
    reference forms them (Vent_Analysis.py:166,:251,:252,:257,:270,:356) from the integer
    counts the kernels return."""
    n = len(stats)
    out = np.zeros((n, len(METRICS)), dtype=np.float64)
    for i in range(n):
        st = stats[i]
        vox = voxs[i]
        nm = np.float64(st["n_mask"])
        voxvol = np.prod(np.divide(vox, 10))
        out[i, 0] = 100 * np.float64(st["n_defect"]) / nm
        out[i, 1] = 100 * np.int64(st["n_lb12"]) / nm
        t95 = int(st["ci_t95"])
        out[i, 2] = ctx.ci_radii(int(handles[i]))[t95] * np.min(vox) if (handles is not None and t95 >= 0) else np.nan
        out[i, 3] = st["snr"]
        out[i, 4] = np.int64(st["n_mask"] if n_mask_eq1 is None else n_mask_eq1[i]) * voxvol / 1000
        out[i, 5] = np.int64(st["n_defect"]) * voxvol / 1000
    return out


def run_cohort_host(images, masks, voxs, device=0, flags=ALL_STAGES, thresh=0.6, ctx=None, params=None,
                    metrics_only=False):
    """Whole path on a list of subjects given as host arrays.  Returns
    ``(metrics [n,6], outputs dict of concatenated arrays, stats)``.  ``metrics_only``: cohort
    screening -- only the per-subject records come back from the GPU (no output arrays)."""
    ctx = ctx if ctx is not None else _lib.context(device)
    shapes = [tuple(int(s) for s in im.shape) for im in images]
    img = np.concatenate([np.asarray(im, dtype=np.float32).ravel() for im in images])
    msk = np.concatenate([np.asarray(m).astype(np.float32).astype(np.uint8).ravel() for m in masks])
    handles = table_handles(ctx, shapes, voxs) if flags & _lib.DO_CI else None
    out = ctx.pipeline_host(shapes, flags, image=img, mask=msk, handles=handles, thresh=thresh, params=params,
                            stats_only=metrics_only)
    if out["rc"] == _lib.VB_ERR_CI_RMAX:
        raise ValueError("--MAX RADIUS REACHED-- for at least one subject (CI.py:101-103)")
    return metrics_from_stats(ctx, out["stats"], shapes, voxs, handles), out, out["stats"]


def build4DdataArrays(images, masks, voxs, out, protons=None, device=0, ctx=None, rmax=50):
    """``Vent_Analysis.build4DdataArray`` (Vent_Analysis.py:292-313) for a cohort, packed on the GPU
    from the compact outputs of :func:`run_cohort_host` (``out``: n4 f32, defect u8, ci_shell u16).
    Returns one float32 ``[rows, cols, slices, 6]`` array per subject: proton[0] HPvent[1] mask[2]
    N4HPvent[3] defectArray[4] CIarray[5]."""
    ctx = ctx if ctx is not None else _lib.context(device)
    shapes = [tuple(int(s) for s in im.shape) for im in images]
    handles = table_handles(ctx, shapes, voxs, rmax) if out.get("ci_shell") is not None else None
    cat = lambda xs, dt: np.concatenate([np.asarray(x, dtype=dt).ravel() for x in xs])
    packed = ctx.pack6_host(shapes, [min(float(v) for v in vox) for vox in voxs],
                            proton=None if protons is None else cat(protons, np.float32),
                            image=cat(images, np.float32), mask=cat(masks, np.uint8), n4=out.get("n4"),
                            defect=out.get("defect"), ci_shell=out.get("ci_shell"), handles=handles)
    res, off = [], 0
    for shp in shapes:
        V = int(np.prod(shp))
        res.append(packed[6 * off:6 * (off + V)].reshape(shp + (6,)))
        off += V
    return res


class DeviceCohort:
    """A batch resident in HBM.  ``step()`` enqueues the whole N4 -> VDP -> CI path on torch's
    current CUDA stream (no host synchronisation); ``stats()`` copies the per-subject record
    back.  Used by bench.py for the HBM-resident ``value`` and by the multi-GPU runner."""

    def __init__(self, images, masks, voxs, device=0, flags=ALL_STAGES, thresh=0.6, want_logbias=False):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.ctx = _lib.context(device)
        self.shapes = np.ascontiguousarray([tuple(int(s) for s in im.shape) for im in images], dtype=np.int32)
        self.voxs = [list(map(float, v)) for v in voxs]
        self.n = len(images)
        self.flags = int(flags)
        self.thresh = float(thresh)
        self.handles = table_handles(self.ctx, self.shapes, self.voxs) if flags & _lib.DO_CI else None
        img = np.concatenate([np.asarray(im, dtype=np.float32).ravel() for im in images])
        msk = np.concatenate([np.asarray(m).astype(np.float32).astype(np.uint8).ravel() for m in masks])
        self.V = img.size
        with torch.cuda.device(self.dev):
            self.image = torch.from_numpy(img).to(self.dev)
            self.mask = torch.from_numpy(msk).to(self.dev)
            self.n4 = torch.empty(self.V, dtype=torch.float32, device=self.dev)
            self.logbias = torch.empty(self.V, dtype=torch.float32, device=self.dev) if want_logbias else None
            self.defect = torch.empty(self.V, dtype=torch.uint8, device=self.dev)
            self.defect_lb = torch.empty(self.V, dtype=torch.uint8, device=self.dev)
            self.defect_border = torch.empty(self.V, dtype=torch.uint8, device=self.dev)
            self.ci_shell = torch.empty(self.V, dtype=torch.int16, device=self.dev)
            self.stats_dev = torch.zeros(self.n * _lib.STATS_DTYPE.itemsize, dtype=torch.uint8, device=self.dev)
        self.params = self.ctx.default_n4_params()
        b = _lib.Buffers()
        b.image = self.image.data_ptr(); b.mask = self.mask.data_ptr(); b.n4 = self.n4.data_ptr()
        b.logbias = self.logbias.data_ptr() if self.logbias is not None else None
        b.defect = self.defect.data_ptr(); b.defect_lb = self.defect_lb.data_ptr()
        b.defect_border = self.defect_border.data_ptr(); b.ci_shell = self.ci_shell.data_ptr()
        b.stats = self.stats_dev.data_ptr()
        self.buffers = b

    def step(self):
        stream = self.torch.cuda.current_stream(self.dev).cuda_stream
        hd = None if self.handles is None else C.c_void_p(self.handles.ctypes.data)
        rc = self.ctx.lib.vb_pipeline_batch_dev(self.ctx.h, self.n, C.c_void_p(self.shapes.ctypes.data), hd,
                                                C.byref(self.params), self.thresh, self.flags,
                                                C.byref(self.buffers), C.c_void_p(stream))
        self.ctx.check(rc)

    def stats(self):
        raw = self.stats_dev.cpu().numpy()
        return raw.view(_lib.STATS_DTYPE).copy()

    def metrics(self):
        return metrics_from_stats(self.ctx, self.stats(), self.shapes, self.voxs, self.handles)


def gather_metrics(local, n_total, rank, world):
    """All ranks' ``[n_local, 6]`` blocks -> ``[n_total, 6]`` in subject order (round-robin
    sharding).  Uses ``torch.distributed.all_gather`` when a process group is up (NCCL on
    GPUs, gloo in the CPU tests); a single process returns its block."""
    import torch
    import torch.distributed as dist
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return np.asarray(local, dtype=np.float64)
    per = (n_total + world - 1) // world
    buf = np.full((per, len(METRICS)), np.nan, dtype=np.float64)
    buf[:len(local)] = local
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(buf).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = np.full((n_total, len(METRICS)), np.nan, dtype=np.float64)
    for r in range(world):
        idx = shard_indices(n_total, r, world)
        out[idx] = parts[r].cpu().numpy()[:len(idx)]
    return out
