"""ctypes binding of ``libventb200.so`` (C ABI in ``include/ventb200.h``).

The product only ever opens the CUDA library built by ``vent_analysis_b200/build.py``.
There is NO CPU fallback: if the library is missing, or no CUDA device is present,
:func:`load` / :class:`Context` raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libventb200.so")

VB_OK, VB_ERR_CUDA, VB_ERR_ARG, VB_ERR_UNSUPPORTED, VB_ERR_CI_RMAX, VB_ERR_EMPTY, VB_ERR_CI_COLOR = 0, 1, 2, 3, 4, 5, 6
NORM_F32, NORM_F64 = 0, 1
DO_SNR, DO_N4, DO_VDP, DO_CI, DO_MASK_BORDER, HOST_STATS_ONLY = 1, 2, 4, 8, 16, 32
MAX_LEVELS = 8
CI_OFF = 0xFFFF


class N4Params(C.Structure):
    """``vb_n4_params`` -- SimpleITK ``N4BiasFieldCorrectionImageFilter`` settings."""
    _fields_ = [("n_levels", C.c_int32), ("max_iter", C.c_int32 * MAX_LEVELS), ("conv_threshold", C.c_float),
                ("fwhm", C.c_float), ("wiener_noise", C.c_float), ("n_bins", C.c_int32), ("ncp", C.c_int32 * 3),
                ("mask_label", C.c_int32), ("max_total_iters", C.c_int32)]


class SubjectStats(C.Structure):
    """``vb_subject_stats``"""
    _fields_ = [("n4_iters", C.c_int32 * MAX_LEVELS), ("n4_last_conv", C.c_float), ("n4_n_mask", C.c_int32),
                ("n_mask", C.c_int32), ("mean", C.c_float), ("p99", C.c_float), ("n_defect", C.c_int32),
                ("n_lb12", C.c_int32), ("n_noise", C.c_int32), ("signal_mean", C.c_double),
                ("noise_mean", C.c_double), ("noise_std", C.c_double), ("snr", C.c_double), ("ci_n", C.c_int32),
                ("ci_t95", C.c_int32), ("ci_error", C.c_int32), ("reserved", C.c_int32)]


STATS_DTYPE = np.dtype([("n4_iters", np.int32, (MAX_LEVELS,)), ("n4_last_conv", np.float32), ("n4_n_mask", np.int32),
                        ("n_mask", np.int32), ("mean", np.float32), ("p99", np.float32), ("n_defect", np.int32),
                        ("n_lb12", np.int32), ("n_noise", np.int32), ("signal_mean", np.float64),
                        ("noise_mean", np.float64), ("noise_std", np.float64), ("snr", np.float64),
                        ("ci_n", np.int32), ("ci_t95", np.int32), ("ci_error", np.int32), ("reserved", np.int32)],
                       align=True)
assert STATS_DTYPE.itemsize == C.sizeof(SubjectStats), (STATS_DTYPE.itemsize, C.sizeof(SubjectStats))


class Buffers(C.Structure):
    """``vb_buffers``"""
    _fields_ = [("image", C.c_void_p), ("mask", C.c_void_p), ("n4", C.c_void_p), ("logbias", C.c_void_p),
                ("defect", C.c_void_p), ("defect_lb", C.c_void_p), ("defect_border", C.c_void_p),
                ("mask_border", C.c_void_p), ("ci_shell", C.c_void_p), ("stats", C.c_void_p)]


class HostIO(C.Structure):
    """``vb_host_io`` -- compact host-side I/O of ``vb_pipeline_batch_host_packed``"""
    _fields_ = [("image_u16", C.c_void_p), ("image_f32", C.c_void_p), ("mask_bits", C.c_void_p), ("mask_u8", C.c_void_p),
                ("n4", C.c_void_p), ("logbias", C.c_void_p), ("codes", C.c_void_p), ("ci_index", C.c_void_p),
                ("ci_value", C.c_void_p), ("ci_capacity", C.c_int64), ("ci_offset", C.c_void_p), ("stats", C.c_void_p)]


class MontageInputs(C.Structure):
    """``vb_montage_inputs``"""
    _fields_ = [("proton", C.c_void_p), ("image", C.c_void_p), ("n4", C.c_void_p), ("mask", C.c_void_p),
                ("mask_border", C.c_void_p), ("defect", C.c_void_p), ("ci_shell", C.c_void_p),
                ("image_mode", C.c_int32), ("proton_mode", C.c_int32)]


CODE_LB_MASK, CODE_DEFECT, CODE_DEFECT_BORDER, CODE_MASK_BORDER = 7, 8, 16, 32


def pack_mask_bits(masks):
    """The ``mask_bits`` layout of ``vb_host_io``: per subject ``ceil(V/32)`` little-endian 32-bit
    words, bit b of word w = voxel ``32 w + b`` of the C-ordered 0/1 volume; subjects concatenated."""
    parts = []
    for m in masks:
        b = np.packbits(np.asarray(m).ravel() != 0, bitorder="little")
        pad = (-b.size) % 4
        parts.append(np.concatenate([b, np.zeros(pad, np.uint8)]) if pad else b)
    return np.ascontiguousarray(np.concatenate(parts)).view(np.uint32)


def mask_words(shapes):
    """Number of ``mask_bits`` words of a batch."""
    return int(sum((int(np.prod(s)) + 31) // 32 for s in shapes))


class VbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ventb200 error {code}: {msg}")
        self.code = code


EXPORTS = ["vb_version", "vb_last_error", "vb_device_count", "vb_create", "vb_destroy", "vb_launch_count",
           "vb_n4_default_params", "vb_ci_table_create", "vb_ci_table_radii", "vb_snr_batch_dev", "vb_n4_batch_dev",
           "vb_vdp_batch_dev", "vb_border_batch_dev", "vb_ci_batch_dev", "vb_pipeline_batch_dev",
           "vb_pipeline_batch_host", "vb_pipeline_batch_host_packed", "vb_host_alloc", "vb_host_free", "vb_overlay_rgb_batch_dev", "vb_overlay_rgb_batch_host", "vb_crop_boxes_batch_host",
           "vb_montage_batch_host", "vb_recon_batch_dev", "vb_recon_batch_host", "vb_pack6_batch_dev", "vb_pack6_batch_host", "vb_set_profiling", "vb_profile_fetch", "vb_set_debug", "vb_n4_debug_fetch",
           "vb_set_n4_cluster", "vb_n4_cluster_size"]
N_STAGES = 5
STAGE_NAMES = ("snr", "n4", "vdp", "mask_border", "ci")


def bind(path):
    """Open a library exporting the ventb200 C ABI and declare its prototypes."""
    lib = C.CDLL(path)
    vp, i32, i64, u32, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float
    lib.vb_version.restype = C.c_int
    lib.vb_last_error.restype = C.c_char_p
    lib.vb_device_count.restype = C.c_int
    lib.vb_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.vb_destroy.argtypes = [vp]
    lib.vb_destroy.restype = None
    lib.vb_launch_count.argtypes = [vp]
    lib.vb_launch_count.restype = C.c_uint64
    lib.vb_n4_default_params.argtypes = [C.POINTER(N4Params)]
    lib.vb_n4_default_params.restype = None
    lib.vb_ci_table_create.argtypes = [vp, vp, i64, i32, i32, i32, C.POINTER(i32)]
    lib.vb_ci_table_radii.argtypes = [vp, i32, vp, i32, C.POINTER(i32)]
    lib.vb_snr_batch_dev.argtypes = [vp, vp, vp, i32, vp, i32, vp, vp]
    lib.vb_n4_batch_dev.argtypes = [vp, vp, vp, i32, vp, C.POINTER(N4Params), vp, vp, vp, vp]
    lib.vb_vdp_batch_dev.argtypes = [vp, vp, vp, i32, vp, f32, vp, vp, vp, vp, vp]
    lib.vb_border_batch_dev.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.vb_ci_batch_dev.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    lib.vb_pipeline_batch_dev.argtypes = [vp, i32, vp, vp, C.POINTER(N4Params), f32, u32, C.POINTER(Buffers), vp]
    lib.vb_pipeline_batch_host.argtypes = [vp, i32, vp, vp, C.POINTER(N4Params), f32, u32, C.POINTER(Buffers)]
    lib.vb_pipeline_batch_host_packed.argtypes = [vp, i32, vp, vp, C.POINTER(N4Params), f32, u32, C.POINTER(HostIO)]
    lib.vb_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.vb_host_free.argtypes = [vp, vp]
    lib.vb_overlay_rgb_batch_dev.argtypes = [vp, i32, vp, vp, vp, i32, i32, vp, vp]
    lib.vb_overlay_rgb_batch_host.argtypes = [vp, i32, vp, vp, vp, i32, i32, vp]
    lib.vb_crop_boxes_batch_host.argtypes = [vp, i32, vp, vp, i32, i32, vp]
    lib.vb_montage_batch_host.argtypes = [vp, i32, vp, vp, vp, C.POINTER(MontageInputs), vp, vp, vp, vp, i64]
    lib.vb_recon_batch_dev.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.vb_recon_batch_host.argtypes = [vp, i32, vp, vp, vp]
    lib.vb_pack6_batch_dev.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.vb_pack6_batch_host.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.vb_set_profiling.argtypes = [vp, i32]
    lib.vb_profile_fetch.argtypes = [vp, vp, C.POINTER(i32)]
    lib.vb_set_debug.argtypes = [vp, i32]
    lib.vb_set_n4_cluster.argtypes = [vp, i32]
    lib.vb_n4_cluster_size.argtypes = [vp]
    lib.vb_n4_cluster_size.restype = i32
    lib.vb_n4_debug_fetch.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, i32]
    return lib


_LIB = None


def load():
    """The CUDA library (and only that).  Raises if it has not been built."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("VENTB200_LIB", LIB_PATH)     # tuning builds only; same C ABI, same CUDA sources
        if path != LIB_PATH:
            _LIB = bind(path)
            return _LIB
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m vent_analysis_b200.build` "
                "(nvcc, sm_100a).  vent_analysis_b200 has no CPU fallback.")
        _LIB = bind(LIB_PATH)
    return _LIB


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """One ``vb_ctx`` (one CUDA device).  ``lib`` is only overridden by the test-suite's
    emulation harness (tests/emu)."""

    def __init__(self, device=0, lib=None):
        self.lib = lib if lib is not None else load()
        h = C.c_void_p()
        rc = self.lib.vb_create(int(device), C.byref(h))
        if rc != VB_OK:
            raise VbError(rc, self.lib.vb_last_error().decode())
        self.h = h
        self.device = int(device)
        self._tables = {}
        self._radii = {}
        self._pinned = []

    def close(self):
        if getattr(self, "h", None):
            for p in self._pinned:
                self.lib.vb_host_free(self.h, p)
            self._pinned = []
            self.lib.vb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != VB_OK:
            raise VbError(rc, self.lib.vb_last_error().decode())

    def launch_count(self):
        return int(self.lib.vb_launch_count(self.h))

    def default_n4_params(self, **kw):
        p = N4Params()
        self.lib.vb_n4_default_params(C.byref(p))
        if "max_iter" in kw:
            mi = list(kw.pop("max_iter"))
            p.n_levels = len(mi)
            for i, v in enumerate(mi):
                p.max_iter[i] = int(v)
        if "ncp" in kw:
            for i, v in enumerate(kw.pop("ncp")):
                p.ncp[i] = int(v)
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    # -- CI tables ---------------------------------------------------------------------------
    def ci_table(self, table, shape, key=None):
        """Handle of the sphere table prepared for volume ``shape`` (memoised by ``key``)."""
        shape = tuple(int(s) for s in shape)
        t = None
        if key is None:                                  # content hash: an id() can be recycled
            import hashlib
            t = np.ascontiguousarray(table, dtype=np.float64)
            key = (t.shape, hashlib.sha1(t.tobytes()).hexdigest())
        k = (key, shape)
        if k not in self._tables:
            t = np.ascontiguousarray(table, dtype=np.float64) if t is None else t
            h = C.c_int32(-1)
            self.check(self.lib.vb_ci_table_create(self.h, _ptr(t), t.shape[0], shape[0], shape[1], shape[2],
                                                   C.byref(h)))
            nb = C.c_int32(0)
            self.check(self.lib.vb_ci_table_radii(self.h, h.value, None, 0, C.byref(nb)))
            radii = np.zeros(nb.value, dtype=np.float64)
            self.check(self.lib.vb_ci_table_radii(self.h, h.value, _ptr(radii), nb.value, C.byref(nb)))
            self._tables[k] = h.value
            self._radii[h.value] = radii
        return self._tables[k]

    def ci_radii(self, handle):
        return self._radii[handle]

    # -- whole path, host buffers ----------------------------------------------------------------
    def pipeline_host(self, shapes, flags, image=None, mask=None, n4=None, handles=None, params=None, thresh=0.6,
                      want_logbias=False, want_mask_border=False, buffers=None, stats_only=False):
        """Run the selected stages on a batch given as HOST arrays (concatenated, C order).
        Returns a dict of numpy outputs (compact dtypes) and the per-subject stats record array.
        ``buffers``: optional dict of preallocated (e.g. pinned) output arrays to fill instead of
        allocating new ones (keys as in the returned dict)."""
        def _buf(key, dtype, size=None):
            size = V if size is None else size
            if buffers is not None and key in buffers:
                a = buffers[key]
                assert a.dtype == dtype and a.size == size and a.flags.c_contiguous, key
                return a.reshape(-1)
            return np.empty(size, dtype=dtype)

        shapes = np.ascontiguousarray(shapes, dtype=np.int32).reshape(-1, 3)
        n = shapes.shape[0]
        V = int(np.prod(shapes.astype(np.int64), axis=1).sum())
        mask = np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1)
        assert mask.size == V
        out = {}
        b = Buffers()
        b.mask = mask.ctypes.data
        if stats_only:                                   # per-subject records only (no output arrays)
            if image is not None:
                image = np.ascontiguousarray(image, dtype=np.float32).reshape(-1)
                b.image = image.ctypes.data
            stats = np.zeros(n, dtype=STATS_DTYPE)
            b.stats = stats.ctypes.data
            hd = None if handles is None else np.ascontiguousarray(handles, dtype=np.int32)
            p = params if params is not None else self.default_n4_params()
            rc = self.lib.vb_pipeline_batch_host(self.h, n, _ptr(shapes), _ptr(hd), C.byref(p), float(thresh),
                                                 int(flags) | HOST_STATS_ONLY, C.byref(b))
            if rc not in (VB_OK, VB_ERR_CI_RMAX):
                self.check(rc)
            return {"stats": stats, "rc": rc}
        if image is not None:
            image = np.ascontiguousarray(image, dtype=np.float32).reshape(-1)
            assert image.size == V
            b.image = image.ctypes.data
        if flags & (DO_N4 | DO_VDP):
            if flags & DO_N4:
                out["n4"] = _buf("n4", np.float32)
            else:
                out["n4"] = np.ascontiguousarray(n4, dtype=np.float32).reshape(-1)
            b.n4 = out["n4"].ctypes.data
        if want_logbias and flags & DO_N4:
            out["logbias"] = _buf("logbias", np.float32)
            b.logbias = out["logbias"].ctypes.data
        if flags & DO_VDP:
            for k in ("defect", "defect_lb", "defect_border"):
                out[k] = _buf(k, np.uint8)
                setattr(b, k, out[k].ctypes.data)
        elif flags & DO_CI:
            raise ValueError("CI without VDP: use ci_host()")
        if want_mask_border or flags & DO_MASK_BORDER:
            flags |= DO_MASK_BORDER
            out["mask_border"] = _buf("mask_border", np.uint8)
            b.mask_border = out["mask_border"].ctypes.data
        if flags & DO_CI:
            out["ci_shell"] = _buf("ci_shell", np.uint16)
            b.ci_shell = out["ci_shell"].ctypes.data
        stats = _buf("stats", STATS_DTYPE, n) if buffers is not None and "stats" in buffers else np.zeros(n, dtype=STATS_DTYPE)
        b.stats = stats.ctypes.data
        hd = None if handles is None else np.ascontiguousarray(handles, dtype=np.int32)
        p = params if params is not None else self.default_n4_params()
        rc = self.lib.vb_pipeline_batch_host(self.h, n, _ptr(shapes), _ptr(hd), C.byref(p), float(thresh),
                                             int(flags), C.byref(b))
        out["stats"] = stats
        out["rc"] = rc
        if rc not in (VB_OK, VB_ERR_CI_RMAX):
            self.check(rc)
        return out

    # -- whole path, compact host I/O --------------------------------------------------------------
    def pinned_empty(self, n, dtype):
        """A page-locked numpy array (``vb_host_alloc``); freed with the context."""
        dtype = np.dtype(dtype)
        p = C.c_void_p()
        nbytes = max(1, int(n) * dtype.itemsize)
        self.check(self.lib.vb_host_alloc(self.h, nbytes, C.byref(p)))
        self._pinned.append(p)
        buf = (C.c_char * nbytes).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(n))

    def pipeline_host_packed(self, shapes, flags, image_u16=None, image_f32=None, mask_bits=None, mask_u8=None,
                             handles=None, params=None, thresh=0.6, want_n4=True, want_logbias=False, want_codes=True,
                             want_ci_map=True, buffers=None):
        """``vb_pipeline_batch_host_packed``: the path on HOST arrays in their compact forms (uint16
        or float32 image, bit-packed or byte mask in; N4HPvent, one code byte per voxel and the
        sparse CI list out).  Returns the raw compact outputs; :func:`unpack_outputs` expands them
        into the arrays of the reference's class surface.  ``buffers``: optional preallocated
        (pinned) output arrays ``n4 / logbias / codes / ci_index / ci_value / ci_offset / stats``."""
        shapes = np.ascontiguousarray(shapes, dtype=np.int32).reshape(-1, 3)
        n = shapes.shape[0]
        V = int(np.prod(shapes.astype(np.int64), axis=1).sum())
        buffers = buffers or {}

        def _buf(key, dtype, size):
            if key in buffers:
                a = buffers[key]
                assert a.dtype == dtype and a.size >= size and a.flags.c_contiguous, key
                return a.reshape(-1)
            return np.empty(size, dtype=dtype)
        io = HostIO()
        keep = []
        for name, arr, dt, size in (("image_u16", image_u16, np.uint16, V), ("image_f32", image_f32, np.float32, V),
                                    ("mask_bits", mask_bits, np.uint32, mask_words(shapes)), ("mask_u8", mask_u8, np.uint8, V)):
            if arr is not None:
                a = np.ascontiguousarray(arr, dtype=dt).reshape(-1)
                assert a.size == size, (name, a.size, size)
                keep.append(a)
                setattr(io, name, a.ctypes.data)
        out = {}
        if want_n4 and flags & DO_N4:
            out["n4"] = _buf("n4", np.float32, V); io.n4 = out["n4"].ctypes.data
        if want_logbias and flags & DO_N4:
            out["logbias"] = _buf("logbias", np.float32, V); io.logbias = out["logbias"].ctypes.data
        if want_codes and flags & DO_VDP:
            out["codes"] = _buf("codes", np.uint8, V); io.codes = out["codes"].ctypes.data
        if want_ci_map and flags & DO_CI:
            cap = buffers["ci_index"].size if "ci_index" in buffers else V
            out["ci_index"] = _buf("ci_index", np.uint32, cap); out["ci_value"] = _buf("ci_value", np.uint16, cap)
            out["ci_offset"] = _buf("ci_offset", np.int64, n + 1)
            io.ci_index = out["ci_index"].ctypes.data; io.ci_value = out["ci_value"].ctypes.data
            io.ci_capacity = cap; io.ci_offset = out["ci_offset"].ctypes.data
        stats = _buf("stats", STATS_DTYPE, n)
        io.stats = stats.ctypes.data
        hd = None if handles is None else np.ascontiguousarray(handles, dtype=np.int32)
        p = params if params is not None else self.default_n4_params()
        rc = self.lib.vb_pipeline_batch_host_packed(self.h, n, _ptr(shapes), _ptr(hd), C.byref(p), float(thresh),
                                                    int(flags), C.byref(io))
        out["stats"] = stats[:n]
        out["rc"] = rc
        if rc not in (VB_OK, VB_ERR_CI_RMAX):
            self.check(rc)
        return out

    # -- overlay compositions ------------------------------------------------------
    @staticmethod
    def norm_mode(a):
        """How the reference's ``normalize`` would compute on array ``a``: float32 arithmetic for a
        float32 array, float64 otherwise (the values must then be exact in float32: the kernels take
        float32 arrays)."""
        a = np.asarray(a)
        if a.dtype == np.float32:
            return NORM_F32, a
        a32 = a.astype(np.float32)
        if not np.array_equal(a32.astype(np.float64), a.astype(np.float64)):
            raise ValueError("overlay: values of a non-float32 array must be exactly representable in float32")
        return NORM_F64, a32

    def overlay_rgb_host(self, shapes, a, b, gui=False):
        """``exportDICOM`` (gui=False, Vent_Analysis.py:387-391) / GUI ``colorBinary`` (gui=True, :628-634)
        overlay of the 0/1 array ``b`` on ``a`` -> uint8 ``[3 V]`` (per subject ``[rows, cols, slices, 3]``)."""
        shapes = np.ascontiguousarray(shapes, dtype=np.int32).reshape(-1, 3)
        V = int(np.prod(shapes.astype(np.int64), axis=1).sum())
        mode, a32 = self.norm_mode(a)
        a32 = np.ascontiguousarray(a32).reshape(-1)
        b8 = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1)
        assert a32.size == V and b8.size == V
        rgb = np.empty(3 * V, dtype=np.uint8)
        self.check(self.lib.vb_overlay_rgb_batch_host(self.h, shapes.shape[0], _ptr(shapes), _ptr(a32), _ptr(b8),
                                                      1 if gui else 0, mode, _ptr(rgb)))
        return rgb

    def crop_boxes_host(self, shapes, mask, border=0, border_slices=False):
        """``cropToData`` boxes (Vent_Analysis.py:430-456): int32 ``[n, 6]`` = r0, r1, c0, c1, s0, s1.
        Raises IndexError like the reference when a subject has no box."""
        shapes = np.ascontiguousarray(shapes, dtype=np.int32).reshape(-1, 3)
        m8 = np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1)
        boxes = np.empty((shapes.shape[0], 6), dtype=np.int32)
        rc = self.lib.vb_crop_boxes_batch_host(self.h, shapes.shape[0], _ptr(shapes), _ptr(m8), int(border),
                                               1 if border_slices else 0, _ptr(boxes))
        if rc == VB_ERR_EMPTY:
            raise IndexError("list index out of range")
        self.check(rc)
        return boxes

    def montage_host(self, shapes, image, n4, mask, mask_border, defect, parula, proton=None, ci_shell=None, handles=None,
                     vox_min=None):
        """``screenShot`` montage (Vent_Analysis.py:466-494, before the text) for a batch: returns
        ``(list of uint8 [7 R, S C, 3] arrays, boxes [n, 6])``.  IndexError like the reference when the
        mask has no crop box or a CI value reaches 40 mm."""
        shapes = np.ascontiguousarray(shapes, dtype=np.int32).reshape(-1, 3)
        n = shapes.shape[0]
        V = int(np.prod(shapes.astype(np.int64), axis=1).sum())
        imode, im32 = self.norm_mode(image)
        inp = MontageInputs()
        keep = [np.ascontiguousarray(im32).reshape(-1), np.ascontiguousarray(n4, dtype=np.float32).reshape(-1),
                np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1), np.ascontiguousarray(mask_border, dtype=np.uint8).reshape(-1),
                np.ascontiguousarray(defect, dtype=np.uint8).reshape(-1)]
        assert all(k.size == V for k in keep)
        inp.image, inp.n4, inp.mask, inp.mask_border, inp.defect = (k.ctypes.data for k in keep)
        inp.image_mode = imode
        if proton is not None:
            pmode, p32 = self.norm_mode(proton)
            p32 = np.ascontiguousarray(p32).reshape(-1)
            keep.append(p32)
            inp.proton, inp.proton_mode = p32.ctypes.data, pmode
        hd = vm = None
        if ci_shell is not None:
            sh = np.ascontiguousarray(ci_shell, dtype=np.uint16).reshape(-1)
            keep.append(sh)
            inp.ci_shell = sh.ctypes.data
            hd = np.ascontiguousarray(handles, dtype=np.int32)
            vm = np.ascontiguousarray(vox_min, dtype=np.float64).reshape(-1)
        par = np.ascontiguousarray(parula, dtype=np.float64)
        assert par.shape == (64, 3)
        boxes = np.empty((n, 6), dtype=np.int32)
        off = np.zeros(n + 1, dtype=np.int64)
        out = np.empty(21 * V, dtype=np.uint8)
        rc = self.lib.vb_montage_batch_host(self.h, n, _ptr(shapes), _ptr(hd), _ptr(vm), C.byref(inp), _ptr(par), _ptr(boxes),
                                            _ptr(off), _ptr(out), out.size)
        if rc == VB_ERR_EMPTY:
            raise IndexError("list index out of range")
        if rc == VB_ERR_CI_COLOR:
            raise IndexError("index out of bounds for axis 0 with size 64")
        self.check(rc)
        imgs = []
        for i in range(n):
            R, Cc, S = (int(boxes[i, 1] - boxes[i, 0]), int(boxes[i, 3] - boxes[i, 2]), int(boxes[i, 5] - boxes[i, 4]))
            imgs.append(out[off[i]:off[i + 1]].reshape(7 * R, S * Cc, 3).copy())
        return imgs, boxes

    def recon_host(self, kspaces):
        """``process_RAW`` reconstruction (Vent_Analysis.py:536-540) of a list of complex k-space volumes
        ``[rows, cols, slices]`` -> list of complex64 images ``[cols, rows, slices]``."""
        shapes = np.ascontiguousarray([k.shape for k in kspaces], dtype=np.int32).reshape(-1, 3)
        flat = np.ascontiguousarray(np.concatenate([np.asarray(k, dtype=np.complex64).ravel() for k in kspaces]))
        out = np.empty_like(flat)
        self.check(self.lib.vb_recon_batch_host(self.h, shapes.shape[0], _ptr(shapes), _ptr(flat), _ptr(out)))
        res, off = [], 0
        for n0, n1, n2 in shapes:
            V = int(n0) * int(n1) * int(n2)
            res.append(out[off:off + V].reshape(int(n1), int(n0), int(n2)))
            off += V
        return res

    def ci_host(self, defect, shapes, handles):
        """CI stage alone on host arrays: defect uint8 0/1 -> (ci_shell uint16, stats, rc)."""
        shapes = np.ascontiguousarray(shapes, dtype=np.int32).reshape(-1, 3)
        n = shapes.shape[0]
        V = int(np.prod(shapes.astype(np.int64), axis=1).sum())
        defect = np.ascontiguousarray(defect, dtype=np.uint8).reshape(-1)
        assert defect.size == V
        b = Buffers()
        dummy_mask = defect
        b.mask = dummy_mask.ctypes.data
        b.defect = defect.ctypes.data
        shell = np.empty(V, dtype=np.uint16)
        b.ci_shell = shell.ctypes.data
        stats = np.zeros(n, dtype=STATS_DTYPE)
        b.stats = stats.ctypes.data
        hd = np.ascontiguousarray(handles, dtype=np.int32)
        p = self.default_n4_params()
        rc = self.lib.vb_pipeline_batch_host(self.h, n, _ptr(shapes), _ptr(hd), C.byref(p), 0.6, DO_CI, C.byref(b))
        if rc not in (VB_OK, VB_ERR_CI_RMAX):
            self.check(rc)
        return shell, stats, rc

    def pack6_host(self, shapes, vox_min, proton=None, image=None, mask=None, n4=None, defect=None, ci_shell=None,
                   handles=None, out=None):
        """``Vent_Analysis.build4DdataArray`` (Vent_Analysis.py:292-313) for a batch on host arrays:
        float32 ``[6 * V]`` = per subject ``[rows][cols][slices][6]`` (proton, HPvent, mask, N4HPvent,
        defectArray, CIarray in mm).  Inputs left ``None`` give a zero channel."""
        shapes = np.ascontiguousarray(shapes, dtype=np.int32).reshape(-1, 3)
        n = shapes.shape[0]
        V = int(np.prod(shapes.astype(np.int64), axis=1).sum())

        def arr(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt).reshape(-1)
            assert a.size == V
            return a
        proton, image, n4 = arr(proton, np.float32), arr(image, np.float32), arr(n4, np.float32)
        mask, defect, ci_shell = arr(mask, np.uint8), arr(defect, np.uint8), arr(ci_shell, np.uint16)
        vm = np.ascontiguousarray(vox_min, dtype=np.float64).reshape(-1)
        assert vm.size == n
        hd = None if handles is None else np.ascontiguousarray(handles, dtype=np.int32)
        if out is None:
            out = np.empty(6 * V, dtype=np.float32)
        assert out.dtype == np.float32 and out.size == 6 * V and out.flags.c_contiguous
        self.check(self.lib.vb_pack6_batch_host(self.h, n, _ptr(shapes), _ptr(hd), _ptr(vm), _ptr(proton), _ptr(image),
                                                _ptr(mask), _ptr(n4), _ptr(defect), _ptr(ci_shell), _ptr(out)))
        return out

    def n4_debug(self, subject, nbins, lattice_capacity, trace_capacity=212):
        self.check(self.lib.vb_set_debug(self.h, 1))
        hist = np.zeros(nbins, dtype=np.float32)
        emap = np.zeros(nbins, dtype=np.float32)
        lat = np.zeros(lattice_capacity, dtype=np.float32)
        ncp = np.zeros(3, dtype=np.int32)
        trace = np.zeros(trace_capacity, dtype=np.float32)
        self.check(self.lib.vb_n4_debug_fetch(self.h, int(subject), _ptr(hist), _ptr(emap), _ptr(lat),
                                              lattice_capacity, _ptr(ncp), _ptr(trace), trace_capacity))
        L = int(np.prod(ncp))
        phases = dict(zip(("setup", "hist", "sharpen", "fit", "update", "chain", "finalize", "other",
                           "chain_prefix", "chain_warp", "chain_team", "chain_unused"), (float(x) for x in trace[-12:])))
        tr = trace[:-12]
        return dict(hist=hist, emap=emap, lattice=lat[:L].reshape(tuple(int(x) for x in ncp)).copy(), ncp=ncp,
                    conv=tr[tr >= 0], phase_us=phases)

    def set_profiling(self, on=True):
        self.check(self.lib.vb_set_profiling(self.h, 1 if on else 0))

    def profile_fetch(self):
        """({stage: summed ms}, number of profiled pipeline calls) since the last fetch."""
        ms = np.zeros(N_STAGES, dtype=np.float64)
        n = C.c_int32(0)
        self.check(self.lib.vb_profile_fetch(self.h, _ptr(ms), C.byref(n)))
        return dict(zip(STAGE_NAMES, (float(x) for x in ms))), n.value

    def set_n4_cluster(self, max_k=0):
        """CTAs (SMs) per subject for N4 on small batches: 0 automatic (volumes of >= 200 k voxels,
        n * K <= 7/8 of the SMs), 1 off, 2 / 4 / 8 forced upper bound (n * K <= SM count)."""
        self.check(self.lib.vb_set_n4_cluster(self.h, int(max_k)))

    def n4_cluster_size(self):
        """CTAs per subject of the most recent N4 launch."""
        return int(self.lib.vb_n4_cluster_size(self.h))

    def set_debug(self, on=True):
        self.check(self.lib.vb_set_debug(self.h, 1 if on else 0))


def unpack_outputs(packed, shapes):
    """Expand the compact outputs of :meth:`Context.pipeline_host_packed` into the dense arrays of
    :meth:`Context.pipeline_host` (``defect``, ``defect_lb``, ``defect_border``, ``mask_border``
    uint8, ``ci_shell`` uint16 with ``CI_OFF`` off the defect) -- format conversion only."""
    shapes = np.asarray(shapes, dtype=np.int64).reshape(-1, 3)
    vols = np.prod(shapes, axis=1)
    V = int(vols.sum())
    out = {k: packed[k] for k in ("n4", "logbias", "stats", "rc") if k in packed}
    if "codes" in packed:
        c = packed["codes"][:V]
        out["defect_lb"] = c & CODE_LB_MASK
        out["defect"] = (c >> 3) & 1
        out["defect_border"] = (c >> 4) & 1
        out["mask_border"] = (c >> 5) & 1
    if "ci_index" in packed:
        shell = np.full(V, CI_OFF, dtype=np.uint16)
        off = packed["ci_offset"]
        total = int(off[len(vols)])
        base = np.repeat(np.concatenate([[0], np.cumsum(vols)[:-1]]), np.diff(off[:len(vols) + 1]))
        shell[base + packed["ci_index"][:total].astype(np.int64)] = packed["ci_value"][:total]
        out["ci_shell"] = shell
    return out


_CTX = {}


def context(device=0):
    """Process-wide context per device (created on first use)."""
    if device not in _CTX:
        _CTX[device] = Context(device)
    return _CTX[device]
