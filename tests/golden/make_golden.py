"""Generates the golden fixtures in this directory by running the UNMODIFIED reference
(``/root/reference/Vent_Analysis.py`` and ``/root/reference/CI.py``) in the build
container.  The reference is Python and cannot travel to the GPU box, so its outputs are
committed here as small ``.npz`` files together with this script.

    python tests/golden/make_golden.py          # needs /root/reference

Recipe: seven absent third-party modules are stubbed in ``sys.modules`` so
that ``import Vent_Analysis`` succeeds; the object is built through the reference's own
``pickle_dict={'vox': ...}`` route (the plain array constructor crashes at
Vent_Analysis.py:166); ``N4_bias_correction`` -- SimpleITK, absent -- is replaced on the
instance by a function returning a fixed float32 volume (the oracle's N4 output), which is
stored (or hashed) in the fixture, so everything downstream of N4 is the reference's own
arithmetic.  CI fixtures call ``CI.calculate_CI`` directly with the working directory set
to ``/root/reference`` so that it loads its shipped sphere tables.
"""
import contextlib
import hashlib
import io
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def load_reference():
    for name in ['nibabel', 'pydicom', 'SimpleITK', 'skimage', 'skimage.util', 'tkinter',
                 'tkinter.filedialog', 'mapvbvd', 'matplotlib', 'matplotlib.pyplot']:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['skimage'].util = sys.modules['skimage.util']
    sys.modules['tkinter'].filedialog = sys.modules['tkinter.filedialog']
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.path.insert(0, REF)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import CI
        import Vent_Analysis
    return Vent_Analysis, CI


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def run_reference(VA, img, mask_f64, vox, n4_volume, with_ci, cwd):
    """Reference pipeline with N4 pinned to ``n4_volume``."""
    old = os.getcwd()
    os.chdir(cwd)
    try:
        v = quiet(VA.Vent_Analysis, xenon_array=img, mask_array=mask_f64, pickle_dict={'vox': list(vox)})
        v.N4_bias_correction = lambda a, b: n4_volume
        quiet(v.calculate_VDP)
        if with_ci:
            quiet(v.calculate_CI)
    finally:
        os.chdir(old)
    return v


def main():
    from oracle import n4_oracle
    from vent_analysis_b200.phantom import CONFIGS, phantom
    VA, CI = load_reference()

    # ---- VDP / SNR / border / LB (+ CI for the small ones) ---------------------------------
    cases = [
        # name, shape, vox, seed, store_inputs, with_ci
        ("small_a", (64, 64, 8), [1.5, 1.5, 10.0], 11, True, True),
        ("small_b", (48, 40, 6), [2.0, 2.0, 11.5], 12, True, True),
        ("ragged_c", (50, 70, 5), [1.5, 1.5, 10.0], 13, True, False),
        ("C1", CONFIGS["C1"][0], CONFIGS["C1"][1], CONFIGS["C1"][2], False, True),
        ("C2", CONFIGS["C2"][0], CONFIGS["C2"][1], CONFIGS["C2"][2], False, True),
    ]
    for name, shape, vox, seed, store, with_ci in cases:
        img, mask = phantom(shape, vox, seed)
        n4 = n4_oracle.n4(img, mask)["out"]
        v = run_reference(VA, img, mask.astype(np.float64), vox, n4, with_ci, REF)
        out = dict(
            shape=np.asarray(shape), vox=np.asarray(vox, dtype=np.float64), seed=seed,
            img_sha=sha(img), mask_sha=sha(mask), n4_sha=sha(n4),
            SNR=np.asarray(v.metadata['SNR']), VDP=np.float64(v.metadata['VDP']),
            VDP_lb=np.float64(v.metadata['VDP_lb']), LungVolume=np.float64(v.metadata['LungVolume']),
            DefectVolume=np.float64(v.metadata['DefectVolume']),
            defect_bits=np.packbits(v.defectArray.astype(np.uint8)),
            defectLB=v.defectArrayLB.astype(np.uint8),
            defectBorder_bits=np.packbits(v.defectBorder.astype(np.uint8)),
            maskBorder_bits=np.packbits(v.mask_border.astype(np.uint8)),
            defect_dtype=str(v.defectArray.dtype), lb_dtype=str(v.defectArrayLB.dtype),
            border_dtype=str(v.defectBorder.dtype), mask_border_dtype=str(v.mask_border.dtype),
            n4_dtype=str(v.N4HPvent.dtype), snr_dtype=str(np.asarray(v.metadata['SNR']).dtype),
        )
        if store:
            out.update(img=img, mask=mask, n4=n4)
        if with_ci:
            sel = v.defectArray > 0
            out.update(CI=np.float64(v.metadata['CI']), CI_values=v.CIarray[sel],
                       ci_dtype=str(v.CIarray.dtype))
        np.savez_compressed(os.path.join(HERE, f"vdp_{name}.npz"), **out)
        print(name, "VDP", out["VDP"], "VDP_lb", out["VDP_lb"], "SNR", out["SNR"],
              "D", int(v.defectArray.sum()), "CI", out.get("CI"))

    # ---- CI alone: edge wrap, large radii, aliasing (s0 <= 98), isotropic regenerated table
    rng = np.random.default_rng(7)
    ci_cases = []
    d = np.zeros((20, 24, 6), dtype=np.float64); d[:9, :11, :4] = 1; d[15:, 20:, 5] = 1
    ci_cases.append(("corner_blob", d, [1.5, 1.5, 10.0], 50, REF))
    d = np.zeros((40, 17, 5), dtype=np.float64); d[30:, :, :] = (rng.random((10, 17, 5)) < 0.7)
    d[3, 3, 2] = 1
    ci_cases.append(("edge_slab", d, [2.0, 2.0, 11.5], 50, REF))
    d = np.zeros((30, 26, 9), dtype=np.float64)
    ii, jj, kk = np.ogrid[:30, :26, :9]
    d[((ii - 12) ** 2 + (jj - 13) ** 2 + 4 * (kk - 4) ** 2) <= 36] = 1
    d[25:28, 2:5, 7:9] = 1
    ci_cases.append(("iso_regen", d, np.array([3.0, 3.0, 3.0]), 12, None))
    for name, d, vox, rmax, cwd in ci_cases:
        tmp = tempfile.mkdtemp() if cwd is None else cwd
        old = os.getcwd(); os.chdir(tmp)
        try:
            ci = quiet(CI.calculate_CI, d, vox, rmax) if rmax != 50 else quiet(CI.calculate_CI, d, vox)
        finally:
            os.chdir(old)
        np.savez_compressed(os.path.join(HERE, f"ci_{name}.npz"), defect=d.astype(np.uint8),
                            vox=np.asarray(vox, dtype=np.float64), rmax=rmax, CIarray=ci,
                            ci_dtype=str(ci.dtype))
        print(name, d.shape, int(d.sum()), "CI max", ci.max())


if __name__ == "__main__":
    main()
