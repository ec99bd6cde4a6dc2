"""Drift pins for N4.  The reference's N4 is SimpleITK 2.3.1's filter (Vent_Analysis.py:330-331),
which is neither installed nor installable in the build container or on the GPU box
(profiles/r02_probe_box_simpleitk.txt), so a golden volume from the REAL filter cannot be made here:
N4 parity stays "vs the restatement" (oracle/n4_oracle.c).  What this script pins is that restatement
itself -- per-level iteration counts, convergence trace, final control lattice and a checksum of the
corrected volume for three phantoms -- so that neither the oracle nor the CUDA path can drift
unnoticed between rounds.  If SimpleITK ever becomes available, regenerate with --simpleitk to
replace the pins by outputs of the real filter.

    python tests/golden/make_golden_n4.py [--simpleitk]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = {"small_a": ((40, 36, 6), [1.5, 1.5, 10.0], 5), "C1": None, "C2": None}


def simpleitk_n4(img, mask):
    import SimpleITK as sitk                                           # Vent_Analysis.py:316-334, verbatim semantics
    image = sitk.Cast(sitk.GetImageFromArray(img.astype(np.float32)), sitk.sitkFloat32)
    m = sitk.Cast(sitk.GetImageFromArray(mask.astype(np.float32)), sitk.sitkUInt8)
    f = sitk.N4BiasFieldCorrectionImageFilter()
    out = sitk.GetArrayFromImage(f.Execute(image, m))
    return out, None


def main():
    from oracle import n4_oracle
    from vent_analysis_b200.phantom import CONFIGS, phantom
    real = "--simpleitk" in sys.argv
    out = {"source": np.array("SimpleITK 2.3.1" if real else "restatement oracle/n4_oracle.c (parity unpinned)")}
    for name, spec in CASES.items():
        shape, vox, seed = spec if spec else CONFIGS[name]
        img, mask = phantom(shape, vox, seed)
        r = n4_oracle.n4(img, mask)
        if real:
            vol, _ = simpleitk_n4(img, mask)
            out[f"{name}_n4_f16"] = vol.astype(np.float16)            # compact copy for a 1e-3 comparison
        else:
            vol = r["out"]
        out[f"{name}_shape"] = np.array(shape)
        out[f"{name}_vox"] = np.array(vox)
        out[f"{name}_seed"] = np.array(seed)
        out[f"{name}_iters"] = r["iters"]
        out[f"{name}_conv"] = np.concatenate([r["conv"][l, :r["iters"][l]] for l in range(len(r["iters"]))])
        out[f"{name}_lattice"] = r["lattice"]
        out[f"{name}_sum"] = np.array(vol.astype(np.float64).sum())
        out[f"{name}_samples"] = vol.ravel()[::max(1, vol.size // 4096)][:4096].copy()
        print(name, shape, r["iters"], float(out[f"{name}_sum"]))
    np.savez_compressed(os.path.join(HERE, "n4_pins.npz"), **out)


if __name__ == "__main__":
    main()
