"""Opt-in pin of the N4 restatement: wherever SimpleITK is importable (it is NOT in the build
image, and there is no network), compare oracle/n4_oracle.c with the reference's actual call
(Vent_Analysis.py:322-332).  Skipped otherwise -- which is why DESIGN.md says N4 parity is
unpinned."""
import numpy as np
import pytest

from oracle import n4_oracle
from vent_analysis_b200.phantom import phantom


def test_restatement_matches_simpleitk():
    sitk = pytest.importorskip("SimpleITK")
    img, mask = phantom((64, 64, 8), [1.5, 1.5, 10.0], 11)
    image = sitk.Cast(sitk.GetImageFromArray(img.astype(np.float32)), sitk.sitkFloat32)
    m = sitk.Cast(sitk.GetImageFromArray(mask.astype(np.float32)), sitk.sitkUInt8)
    corrector = sitk.N4BiasFieldCorrectionImageFilter()
    corrector.SetNumberOfThreads(1)
    ref = sitk.GetArrayFromImage(corrector.Execute(image, m))
    got = n4_oracle.n4(img, mask)["out"]
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6)
    assert rel.max() < 1e-3
