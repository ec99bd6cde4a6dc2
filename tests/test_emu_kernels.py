"""CPU suite: the product's KERNEL SOURCES (vent_analysis_b200/csrc), compiled for the host
by tests/emu (one OS thread per CUDA thread), checked against the golden fixtures and the
oracle through the same C ABI.  This pins the kernel logic in the GPU-less build container;
tests/test_gpu_parity.py repeats it on the real sm_100a build."""
import numpy as np
import pytest

import helpers
from emu.harness import emu_context
from oracle import n4_oracle
from vent_analysis_b200 import _lib
from vent_analysis_b200.phantom import phantom


@pytest.fixture(scope="module")
def ctx():
    return emu_context()


@pytest.mark.parametrize("name,with_ci", [("small_a", True), ("small_b", True), ("ragged_c", False)])
def test_vdp_snr_border_ci_bit_exact(ctx, name, with_ci):
    helpers.check_vdp_ci_against_golden(ctx, name, with_ci)


@pytest.mark.parametrize("name", ["corner_blob", "edge_slab", "iso_regen"])
def test_ci_edge_cases_bit_exact(ctx, name):
    helpers.check_ci_case(ctx, name)


def test_ci_table_exhaustion_is_reported(ctx):
    from vent_analysis_b200.sphere_table import generate
    d = np.ones((12, 12, 3), dtype=np.uint8)
    h = ctx.ci_table(generate([1.0, 1.0, 1.0], 2), d.shape, key="exhaust")
    _, stats, rc = ctx.ci_host(d, [d.shape], [h])
    assert rc == _lib.VB_ERR_CI_RMAX and stats["ci_error"][0] == 1


def test_n4_matches_restatement_single_iteration_stages(ctx):
    """One iteration at one level: histogram, E(u|v) map, fitted lattice and convergence value
    against the oracle's stage functions."""
    shape = (40, 36, 6)
    img, mask = phantom(shape, [1.5, 1.5, 10.0], 5)
    ctx.set_debug(True)
    out, ref, _ = helpers.check_n4(ctx, img, mask, max_iter=(1,))
    dbg = ctx.n4_debug(0, 200, 11 ** 3)
    ctx.set_debug(False)
    m = mask > 0
    U = np.where(m, np.log(np.maximum(img, 1e-30)), img).astype(np.float32)
    s = n4_oracle.sharpen(U, m.astype(np.uint8))
    assert abs(dbg["hist"].sum() - m.sum()) < 1e-2
    assert np.abs(dbg["hist"] - s["H"]).max() < 2e-3
    assert np.abs(dbg["emap"] - s["E"]).max() < 1e-5
    assert np.abs(dbg["lattice"] - ref["lattice"]).max() < 1e-6
    assert abs(dbg["conv"][0] - ref["conv"][0][0]) < 1e-6


def test_n4_full_run_matches_restatement(ctx):
    shape = (40, 36, 6)
    img, mask = phantom(shape, [1.5, 1.5, 10.0], 5)
    _, _, rel = helpers.check_n4(ctx, img, mask)
    assert rel < 1e-4


@pytest.fixture(scope="module")
def ctx8(monkeypatch_module=None):
    """An emulated device with 8 SMs: single subjects run with clusters of up to 8 CTAs."""
    import os
    import build_emu
    old = os.environ.get("VB_EMU_SMS")
    os.environ["VB_EMU_SMS"] = "8"
    try:
        c = _lib.Context(0, lib=_lib.bind(os.environ.get("VB_EMU_LIB") or build_emu.build()))
    finally:
        if old is None:
            del os.environ["VB_EMU_SMS"]
        else:
            os.environ["VB_EMU_SMS"] = old
    return c


def _cluster_cases():
    a = phantom((24, 20, 5), [1.5, 1.5, 10.0], 21)
    b = phantom((18, 30, 4), [2.0, 2.0, 11.5], 22)
    e_img, e_m = phantom((20, 16, 3), [2.0, 2.0, 11.5], 23)
    one = np.zeros_like(e_m)
    one[3, 4, 1] = 1
    return [[a], [a, b], [a, (e_img, np.zeros_like(e_m)), (e_img, one), b]]


def test_n4_cluster_of_ctas_per_subject_is_bit_identical(ctx8):
    """K = 2, 4, 8 CTAs per subject (the CTAs of an emulated cluster run concurrently, one forked
    process each): outputs, log bias, iteration counts and the last convergence value equal the
    one-CTA run bit for bit -- single subject, a ragged batch with an empty and a one-voxel mask
    (two ragged subjects: the column-walk test below)."""
    one, _, four = _cluster_cases()
    helpers.check_n4_cluster(ctx8, [one], ks=(4, 8), max_iter=(6, 4, 3), sms=8)
    helpers.check_n4_cluster(ctx8, [four], ks=(2,), max_iter=(4, 3, 2), sms=8)
    # the first two levels run to convergence (the phase team's speculative fit is discarded there, after an odd
    # and an even number of its exchanges: its slot sets must not depend on the chain team's count)
    helpers.check_n4_cluster(ctx8, [[phantom((40, 36, 6), [1.5, 1.5, 10.0], 5)]], ks=(2,), max_iter=(50, 50, 1), sms=8)


def test_n4_cluster_with_the_column_walk_update_and_u_in_global_memory(ctx8, monkeypatch):
    """Same with no shared-memory room for the compact arrays / coefficient tables (the path of large volumes)."""
    monkeypatch.setenv("VB_N4_SMEM_CAP", "0")
    one, two, _ = _cluster_cases()
    helpers.check_n4_cluster(ctx8, [one], ks=(8,), max_iter=(5, 3), sms=8)
    helpers.check_n4_cluster(ctx8, [two], ks=(2,), max_iter=(5, 3), sms=8)


def test_n4_cluster_automatic_choice(ctx8, monkeypatch):
    """vb_set_n4_cluster(0): small volumes keep one CTA per subject; above the size threshold the
    largest K with n * K <= 7/8 of the SMs (8 emulated SMs: 1 subject -> 4 CTAs, 2 or 3 -> 2, 4 -> 1);
    a forced K ignores the threshold and fills the SMs (n * K <= 8)."""
    a = phantom((24, 20, 5), [1.5, 1.5, 10.0], 21)
    p = ctx8.default_n4_params(max_iter=(1,))

    def k_used(n, max_k):
        ctx8.set_n4_cluster(max_k)
        ctx8.pipeline_host([a[0].shape] * n, _lib.DO_N4, image=np.tile(a[0].ravel(), n), mask=np.tile(a[1].ravel(), n), params=p)
        return ctx8.n4_cluster_size()

    try:
        assert k_used(1, 0) == 1                     # 2400 voxels: below the threshold
        assert k_used(1, 8) == 8 and k_used(2, 8) == 4 and k_used(1, 1) == 1
        monkeypatch.setenv("VB_N4_CLUSTER_MIN_V", "0")
        assert [k_used(n, 0) for n in (1, 2, 3, 4)] == [4, 2, 2, 1]
    finally:
        ctx8.set_n4_cluster(0)


def test_n4_ragged_batch_and_label_semantics(ctx):
    """Two subjects of different shapes in one launch; mask label 2 is ignored (MaskLabel=1)."""
    a_img, a_m = phantom((24, 20, 5), [1.5, 1.5, 10.0], 21)
    b_img, b_m = phantom((18, 30, 4), [2.0, 2.0, 11.5], 22)
    b_m = b_m.copy()
    b_m[:2, :, :] = np.where(b_m[:2, :, :] > 0, 2, 0)
    shapes = [a_img.shape, b_img.shape]
    p = ctx.default_n4_params(max_iter=(4, 3))
    out = ctx.pipeline_host(shapes, _lib.DO_N4, image=np.concatenate([a_img.ravel(), b_img.ravel()]),
                            mask=np.concatenate([a_m.ravel(), b_m.ravel()]), params=p, want_logbias=True)
    off = 0
    for img, m in ((a_img, a_m), (b_img, b_m)):
        ref = n4_oracle.n4(img, m, max_iter=(4, 3))
        lb = out["logbias"][off:off + img.size].reshape(img.shape)
        assert np.abs(np.expm1(lb - ref["logbias"])).max() < 1e-4
        off += img.size
    assert list(out["stats"]["n4_iters"][0][:2]) == [4, 3]


def test_edge_cases_ragged_degenerate_inputs(ctx):
    helpers.check_edge_cases(ctx)


def test_class_surface_end_to_end(ctx, monkeypatch):
    """The Vent_Analysis mirror (constructor, calculate_VDP, calculate_CI, metadata, dtypes)
    against the reference's outputs -- N4 here is OUR kernel, so VDP is held to the
    north_star tolerance (+-0.05 pp) and everything downstream of the defect mask to
    bit-exactness given that mask."""
    from vent_analysis_b200 import CI, Vent_Analysis
    from oracle import ci_oracle, vdp_oracle
    from vent_analysis_b200.sphere_table import generate
    monkeypatch.setitem(_lib._CTX, 0, ctx)
    g, shape, vox, img, mask, n4 = helpers.golden_inputs("small_a")
    v = helpers.quiet(Vent_Analysis, xenon_array=img, mask_array=mask.astype(np.float64), vox=vox)
    helpers.quiet(v.calculate_VDP)
    helpers.quiet(v.calculate_CI)
    assert abs(v.metadata["VDP"] - float(g["VDP"])) <= helpers.VDP_PP_TOL
    assert abs(v.metadata["VDP_lb"] - float(g["VDP_lb"])) <= helpers.VDP_PP_TOL
    assert v.metadata["LungVolume"] == float(g["LungVolume"])
    assert np.isclose(v.metadata["SNR"], g["SNR"], rtol=2e-6) and v.metadata["SNR"].dtype == np.float32
    for attr, key in (("defectArray", "defect_dtype"), ("defectArrayLB", "lb_dtype"), ("defectBorder", "border_dtype"),
                      ("mask_border", "mask_border_dtype"), ("N4HPvent", "n4_dtype"), ("CIarray", "ci_dtype")):
        assert getattr(v, attr).dtype == np.dtype(str(g[key])) and getattr(v, attr).shape == shape
    # downstream of OUR defect mask everything is exact vs the oracle
    ref = vdp_oracle.vdp(v.N4HPvent, mask.astype(np.float64), vox)
    assert np.array_equal(v.defectArray, ref["defectArray"])
    assert np.array_equal(v.defectArrayLB, ref["defectArrayLB"])
    assert np.array_equal(v.defectBorder, ref["defectBorder"])
    assert v.metadata["VDP"] == ref["VDP"] and v.metadata["VDP_lb"] == ref["VDP_lb"]
    ci_ref, _ = ci_oracle.calculate_ci(v.defectArray, vox, generate(vox, 50))
    assert np.array_equal(v.CIarray, ci_ref)
    assert v.metadata["CI"] == vdp_oracle.ci_percentile(ci_ref, v.defectArray)
    assert np.array_equal(CI.calculate_CI(v.defectArray, vox), ci_ref)
    # pickling round trip keeps plain numpy state (Vent_Analysis.py:542-559)
    import pickle
    d = {k: val for k, val in vars(v).items()}
    v2 = helpers.quiet(Vent_Analysis, pickle_dict=pickle.loads(pickle.dumps(d)))
    assert np.array_equal(v2.defectArray, v.defectArray) and v2.metadata["CI"] == v.metadata["CI"]
    # reference quirk: the array constructor without vox raises (Vent_Analysis.py:166)
    with pytest.raises(Exception):
        helpers.quiet(Vent_Analysis, xenon_array=img, mask_array=mask.astype(np.float64))


def test_vdp_thresholds_a_few_ulps_around_every_edge(ctx):
    helpers.check_vdp_threshold_edges(ctx)


def test_vdp_p99_select_shared_and_restreamed(ctx):
    helpers.check_p99_select(ctx)


def test_vdp_anchor_mean_is_numpys_pairwise_mean_of_the_sorted_list(ctx):
    helpers.check_mean_bit_exact(ctx, n_phantoms=3)


def test_build4DdataArray_packing_bit_exact(ctx):
    helpers.check_pack6(ctx)


def test_build4DdataArray_identical_to_the_reference_output(ctx):
    helpers.check_pack6_golden(ctx)


def test_compact_host_io_equals_the_dense_path(ctx):
    helpers.check_packed_io(ctx)


def test_overlay_compositions_bit_exact_vs_reference(ctx):
    helpers.check_overlays(ctx)


def test_class_overlay_methods(ctx, monkeypatch, tmp_path):
    """exportDICOM pixel data and the screenShot array of the class mirror equal the reference's own
    (tests/golden/overlay_small_b.npz) when the class is given the fixture's N4 volume."""
    from vent_analysis_b200 import Vent_Analysis
    import types
    monkeypatch.setitem(_lib._CTX, 0, ctx)
    g = helpers.golden("overlay_small_b.npz")
    v0, shape, vox, img, mask, n4 = helpers.golden_inputs("small_b")
    v = helpers.quiet(Vent_Analysis, xenon_array=img, mask_array=mask.astype(np.float64), vox=vox)
    v.N4_bias_correction = lambda a, b: n4                     # (the reference run had N4 pinned the same way)
    out = ctx.pipeline_host([shape], _lib.DO_VDP, mask=mask, n4=n4)
    v.N4HPvent = n4
    v.defectArray = out["defect"].reshape(shape).astype(np.float64)
    v.metadata["VDP"] = 1.0
    helpers.quiet(v.calculate_CI)
    v.proton = (500.0 * np.random.default_rng(int(g["proton_seed"])).random(shape)).astype(np.float32)
    assert np.array_equal(v._overlay_rgb(), g["export_rgb"])
    assert np.array_equal(v.screenShotArray(parula=g["parula"]), g["screenshot"])
    helpers.quiet(v.screenShot, str(tmp_path / "shot.png"), parula=g["parula"])
    assert (tmp_path / "shot.png").stat().st_size > 1000
    assert "defectArray: array(48, 40, 6)" in repr(v)
    # labels other than 0/1: N4 / SNR / borders run like the reference (label 1 / > 0), VDP refuses
    m2 = mask.astype(np.float64).copy(); m2[:24] *= 2
    v2 = helpers.quiet(Vent_Analysis, xenon_array=img, mask_array=m2, vox=vox)
    assert v2.N4_bias_correction(img, m2).shape == shape
    with pytest.raises(ValueError):
        helpers.quiet(v2.calculate_VDP)


def test_kspace_reconstruction_matches_the_reference(ctx):
    helpers.check_recon(ctx)
