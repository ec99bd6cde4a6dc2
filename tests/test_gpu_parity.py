"""GPU suite (B200): the sm_100a library through its C ABI against the golden fixtures
(outputs of the unmodified reference), the oracle, and size-independent properties at
BASELINE.json's full sizes.  `/root/reference` is not read here."""
import numpy as np
import pytest

import helpers
from oracle import ci_oracle, n4_oracle, vdp_oracle
from vent_analysis_b200 import _lib
from vent_analysis_b200.phantom import CONFIGS, cohort_member, phantom
from vent_analysis_b200.sphere_table import generate

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _lib.context(0)


@pytest.mark.parametrize("name,with_ci", [("small_a", True), ("small_b", True), ("ragged_c", False),
                                          ("C1", True), ("C2", True)])
def test_vdp_snr_border_ci_bit_exact_vs_reference(ctx, name, with_ci):
    helpers.check_vdp_ci_against_golden(ctx, name, with_ci)


@pytest.mark.parametrize("name", ["corner_blob", "edge_slab", "iso_regen"])
def test_ci_edge_cases_bit_exact_vs_reference(ctx, name):
    helpers.check_ci_case(ctx, name)


def test_ci_exhaustion_raises_like_reference(ctx):
    from vent_analysis_b200 import CI
    d = np.ones((12, 12, 3), dtype=np.uint8)
    h = ctx.ci_table(generate([1.0, 1.0, 1.0], 2), d.shape, key="exhaust")
    _, stats, rc = ctx.ci_host(d, [d.shape], [h])
    assert rc == _lib.VB_ERR_CI_RMAX and stats["ci_error"][0] == 1
    with pytest.raises(ValueError):
        CI.calculate_CI(np.ones((12, 12, 3)), [1.0, 1.0, 1.0], Rmax=2)
    with pytest.raises(IndexError):
        CI.calculate_CI(np.zeros((12, 12, 3)), [1.0, 1.0, 1.0], Rmax=2)


@pytest.mark.parametrize("shape,vox,seed", [((40, 36, 6), [1.5, 1.5, 10.0], 5), ((64, 64, 8), [1.5, 1.5, 10.0], 11),
                                            (CONFIGS["C1"][0], CONFIGS["C1"][1], 0),
                                            (CONFIGS["C2"][0], CONFIGS["C2"][1], 1)])
def test_n4_bias_field_within_1e3_of_restatement(ctx, shape, vox, seed):
    img, mask = phantom(shape, vox, seed)
    _, _, rel = helpers.check_n4(ctx, img, mask)
    assert rel < helpers.N4_BIAS_RTOL


def test_n4_cluster_of_ctas_per_subject_is_bit_identical(ctx, monkeypatch):
    """Small batches run N4 with a thread-block cluster of K CTAs (K SMs) per subject
    (vb_set_n4_cluster; the reference's own use is one subject at a time): corrected volume, log
    bias, per-level iteration counts and last convergence value equal the one-CTA-per-subject run
    BIT FOR BIT at K = 2, 4, 8 -- one C2 subject, a ragged batch with an empty and a one-voxel mask,
    16 C1 subjects (K <= 8), 30 (K <= 4), and a subject whose compact arrays do not fit shared
    memory (208x256x18); then the same with the column-walk update (no table space)."""
    a = phantom((48, 40, 6), [1.5, 1.5, 10.0], 21)
    b = phantom((33, 50, 4), [2.0, 2.0, 11.5], 22)
    e_img, e_m = phantom((20, 16, 3), [2.0, 2.0, 11.5], 23)
    one = np.zeros_like(e_m)
    one[3, 4, 1] = 1
    c2 = phantom(CONFIGS["C2"][0], CONFIGS["C2"][1], 3)
    c1 = [phantom(CONFIGS["C1"][0], CONFIGS["C1"][1], 40 + i) for i in range(30)]
    sms = 148
    helpers.check_n4_cluster(ctx, [[c2], [a, (e_img, np.zeros_like(e_m)), (e_img, one), b], c1[:16], c1], sms=sms)
    helpers.check_n4_cluster(ctx, [[phantom((208, 256, 18), [1.5, 1.5, 10.0], 9)]], ks=(8,), sms=sms)
    monkeypatch.setenv("VB_N4_SMEM_CAP", "0")
    helpers.check_n4_cluster(ctx, [[c2], [a, b]], ks=(2, 8), sms=sms)


def test_n4_cluster_automatic_choice_on_the_device(ctx, monkeypatch):
    """vb_set_n4_cluster(0) on the 148-SM device: a C2 subject alone gets 8 CTAs; 16 subjects 8 each,
    18 -> 4, 40 -> 2, 74 -> 1 (n * K <= 7/8 of the SMs); a small volume keeps one CTA; the chunks of a
    multi-stream host schedule keep one CTA per subject -- and that run equals the unchunked one."""
    c2 = phantom(CONFIGS["C2"][0], CONFIGS["C2"][1], 3)
    tiny = phantom((48, 40, 6), [1.5, 1.5, 10.0], 21)
    p = ctx.default_n4_params(max_iter=(2, 1))

    def run(sub, n):
        out = ctx.pipeline_host([sub[0].shape] * n, _lib.DO_N4, image=np.tile(sub[0].ravel(), n), mask=np.tile(sub[1].ravel(), n), params=p)
        return ctx.n4_cluster_size(), out["n4"].copy()

    ctx.set_n4_cluster(0)
    assert [run(c2, n)[0] for n in (1, 16, 18, 40, 74)] == [8, 8, 4, 2, 1]
    assert run(tiny, 1)[0] == 1
    k_one, whole = run(c2, 4)
    monkeypatch.setenv("VB_CHUNK_PER", "2")
    k_chunked, chunked = run(c2, 4)
    assert (k_one, k_chunked) == (8, 1) and chunked.tobytes() == whole.tobytes()


def test_n4_stage_level_parity(ctx):
    shape = (40, 36, 6)
    img, mask = phantom(shape, [1.5, 1.5, 10.0], 5)
    ctx.set_debug(True)
    out, ref, _ = helpers.check_n4(ctx, img, mask, max_iter=(1,))
    dbg = ctx.n4_debug(0, 200, 11 ** 3)
    ctx.set_debug(False)
    m = mask > 0
    U = np.where(m, np.log(np.maximum(img, 1e-30)), img).astype(np.float32)
    s = n4_oracle.sharpen(U, m.astype(np.uint8))
    assert np.abs(dbg["hist"] - s["H"]).max() < 2e-3
    assert np.abs(dbg["emap"] - s["E"]).max() < 1e-5
    assert np.abs(dbg["lattice"] - ref["lattice"]).max() < 1e-6


def test_n4_c4_256x256x64_four_levels(ctx):
    """Config 4: high-res volume, multilevel N4 (4 levels) + linear-binning VDP."""
    shape, vox, seed = CONFIGS["C4"]
    img, mask = phantom(shape, vox, seed)
    out, ref, rel = helpers.check_n4(ctx, img, mask, require_same_iters=False)
    r = vdp_oracle.vdp(out["n4"].reshape(shape), mask.astype(np.float64), vox)
    o2 = ctx.pipeline_host([shape], _lib.DO_VDP, mask=mask, n4=out["n4"])
    assert np.array_equal(o2["defect_lb"].reshape(shape), r["defectArrayLB"])
    assert np.array_equal(o2["defect"].reshape(shape), r["defectArray"])
    r_ref = vdp_oracle.vdp(ref["out"], mask.astype(np.float64), vox)
    assert abs(100 * np.int64(o2["stats"][0]["n_lb12"]) / mask.sum() - r_ref["VDP_lb"]) <= helpers.VDP_PP_TOL


def test_ci_c3_isotropic_large_table(ctx):
    """Config 3: 128^3 @3.9 mm, table regenerated for that spacing (523k rows, bitset too large
    for shared memory -> global-memory variant of the kernel).  Checked against the C oracle on
    the full volume (bit exact)."""
    shape, vox, seed = CONFIGS["C3"]
    img, mask = phantom(shape, vox, seed)
    n4 = n4_oracle.n4(img, mask, max_iter=(8, 4))["out"]
    r = vdp_oracle.vdp(n4, mask.astype(np.float64), vox)
    tbl = generate(vox, 50)
    h = ctx.ci_table(tbl, shape, key=(tuple(vox), 50, tbl.shape[0]))
    shell, stats, rc = ctx.ci_host(r["defectArray"], [shape], [h])
    assert rc == 0
    ci_ref, shell_ref = ci_oracle.calculate_ci(r["defectArray"], vox, tbl)
    sel = r["defectArray"] > 0
    assert np.array_equal(shell.reshape(shape)[sel].astype(np.int32), shell_ref[sel])
    assert ctx.ci_radii(h)[int(stats["ci_t95"][0])] * min(vox) == vdp_oracle.ci_percentile(ci_ref, r["defectArray"])


def test_edge_cases_ragged_degenerate_inputs(ctx):
    helpers.check_edge_cases(ctx)


def test_vdp_thresholds_a_few_ulps_around_every_edge(ctx):
    helpers.check_vdp_threshold_edges(ctx)


def test_vdp_p99_select_shared_and_restreamed(ctx):
    helpers.check_p99_select(ctx)


def test_vdp_anchor_mean_is_numpys_pairwise_mean_of_the_sorted_list(ctx):
    """64 C2-sized phantoms + the corner cases + a 128^3 volume (mask beyond shared memory and a
    pairwise heap deeper than the shared-memory one)."""
    shape, vox, seed = CONFIGS["C3"]
    helpers.check_mean_bit_exact(ctx, n_phantoms=64, shape=CONFIGS["C2"][0], big=phantom(shape, vox, seed))


def test_class_surface_end_to_end_c2():
    from vent_analysis_b200 import CI, Vent_Analysis
    g, shape, vox, img, mask, n4 = helpers.golden_inputs("C2")
    v = helpers.quiet(Vent_Analysis, xenon_array=img, mask_array=mask.astype(np.float64), vox=vox)
    helpers.quiet(v.calculate_VDP)
    helpers.quiet(v.calculate_CI)
    # vs the reference run (whose N4 was the oracle restatement): north_star tolerances
    assert abs(v.metadata["VDP"] - float(g["VDP"])) <= helpers.VDP_PP_TOL
    assert abs(v.metadata["VDP_lb"] - float(g["VDP_lb"])) <= helpers.VDP_PP_TOL
    assert v.metadata["LungVolume"] == float(g["LungVolume"])
    assert np.isclose(v.metadata["SNR"], g["SNR"], rtol=2e-6)
    rel = np.abs(v.N4HPvent - n4) / np.maximum(np.abs(n4), 1e-6)
    assert rel.max() < helpers.N4_BIAS_RTOL
    ref = vdp_oracle.vdp(v.N4HPvent, mask.astype(np.float64), vox)
    assert np.array_equal(v.defectArray, ref["defectArray"]) and np.array_equal(v.defectArrayLB, ref["defectArrayLB"])
    ci_ref, _ = ci_oracle.calculate_ci(v.defectArray, vox, generate(vox, 50))
    assert np.array_equal(v.CIarray, ci_ref)
    assert v.metadata["CI"] == vdp_oracle.ci_percentile(ci_ref, v.defectArray)
    if np.array_equal(v.defectArray, helpers.unpack(g["defect_bits"], shape)):
        assert v.metadata["CI"] == float(g["CI"]) and v.metadata["VDP"] == float(g["VDP"])
    assert np.array_equal(v.calculateBorder(v.mask), helpers.unpack(g["maskBorder_bits"], shape))
    assert v.calculate_SNR(v.HPvent).dtype == np.float32


def test_cohort_mixed_spacings_host_and_device_paths_agree(ctx):
    """Config 5 in miniature: 12 subjects, C1/C2 alternating, one launch sequence; per-subject
    results equal the single-subject path bit for bit (batching / sharding changes nothing),
    and the HBM-resident path equals the host-buffer path."""
    from vent_analysis_b200 import cohort
    subs = [cohort_member(i) for i in range(12)]
    imgs, masks, voxs = zip(*subs)
    metrics, out, stats = cohort.run_cohort_host(imgs, masks, voxs)
    dc = cohort.DeviceCohort(imgs, masks, voxs)
    dc.step()
    dc.torch.cuda.synchronize()
    m2 = dc.metrics()
    assert np.array_equal(metrics, m2)
    off = 0
    for i in (0, 5, 11):
        off = sum(im.size for im in imgs[:i])
        one, o1, _ = cohort.run_cohort_host([imgs[i]], [masks[i]], [voxs[i]])
        assert np.array_equal(one[0], metrics[i])
        V = imgs[i].size
        assert np.array_equal(o1["n4"], out["n4"][off:off + V])
        assert np.array_equal(o1["ci_shell"], out["ci_shell"][off:off + V])
    assert np.isfinite(metrics).all()
    # screening mode: only the per-subject records come back, same numbers
    m_only, o_only, _ = cohort.run_cohort_host(imgs, masks, voxs, metrics_only=True)
    assert np.array_equal(m_only, metrics) and set(o_only) == {"stats", "rc"}
    # determinism: same batch twice -> identical bits
    metrics3, out3, _ = cohort.run_cohort_host(imgs, masks, voxs)
    assert np.array_equal(out3["n4"], out["n4"]) and np.array_equal(metrics3, metrics)


def test_properties_at_full_size(ctx):
    """Size-independent properties on C1: N4 ignores a constant gain (bias field unchanged,
    output scales), defect mask and CI invariant under that gain, CI idempotent on repeat,
    CI >= min(vox) everywhere on defects and 0 elsewhere."""
    from vent_analysis_b200 import cohort
    img, mask, vox = cohort_member(0)
    m1, o1, _ = cohort.run_cohort_host([img], [mask], [vox])
    m2, o2, _ = cohort.run_cohort_host([img * 4.0], [mask], [vox])
    assert np.allclose(o2["n4"], 4.0 * o1["n4"], rtol=2e-3)
    assert abs(m1[0, 0] - m2[0, 0]) <= helpers.VDP_PP_TOL
    sel = o1["defect"] > 0
    assert (o1["ci_shell"][~sel] == _lib.CI_OFF).all() and (o1["ci_shell"][sel] < 5000).all()
    m3, o3, _ = cohort.run_cohort_host([img], [mask], [vox])
    assert np.array_equal(o3["ci_shell"], o1["ci_shell"])


def test_build4DdataArray_packing_bit_exact(ctx):
    helpers.check_pack6(ctx)


def test_build4DdataArray_identical_to_the_reference_output(ctx):
    helpers.check_pack6_golden(ctx)


def test_compact_host_io_equals_the_dense_path(ctx):
    helpers.check_packed_io(ctx)


def test_overlay_compositions_bit_exact_vs_reference(ctx):
    helpers.check_overlays(ctx)


def test_kspace_reconstruction_matches_the_reference(ctx):
    helpers.check_recon(ctx)


@pytest.mark.parametrize("name", ["small_a", "C1", "C2"])
def test_n4_against_the_committed_pins(ctx, name):
    """The CUDA N4 against tests/golden/n4_pins.npz: identical per-level iteration counts, final
    control lattice within 1e-4, sampled corrected voxels within the north-star 1e-3.  (The pins are
    outputs of the restatement; SimpleITK is not installable on either box,
    profiles/r02_probe_box_simpleitk.txt.)"""
    g = helpers.golden("n4_pins.npz")
    shape, vox, seed = tuple(int(x) for x in g[f"{name}_shape"]), [float(x) for x in g[f"{name}_vox"]], int(g[f"{name}_seed"])
    img, mask = phantom(shape, vox, seed)
    ctx.set_debug(True)
    out = ctx.pipeline_host([shape], _lib.DO_N4, image=img, mask=mask)
    dbg = ctx.n4_debug(0, 200, 11 ** 3)
    ctx.set_debug(False)
    assert [int(x) for x in out["stats"]["n4_iters"][0][:4]] == [int(x) for x in g[f"{name}_iters"]]
    assert np.abs(dbg["lattice"] - g[f"{name}_lattice"]).max() < 1e-4
    got = out["n4"].ravel()[::max(1, out["n4"].size // 4096)][:4096]
    ref = g[f"{name}_samples"]
    assert (np.abs(got - ref) <= helpers.N4_BIAS_RTOL * np.maximum(np.abs(ref), 1e-6)).all()


def test_stress_slice_64_cohort_phantoms(ctx):
    """A 64-subject slice of the three stress tools (tools/stress_parity.py, stress_vdp_ci.py,
    stress_shapes.py) in ONE batch: N4 iteration counts identical to the restatement and bias field
    within 1e-3 on every subject (oracle on the host cores), VDP / LB / CI bit-exact downstream of
    our N4 on eight of them."""
    import multiprocessing as mp
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import stress_parity
    from vent_analysis_b200 import cohort
    seeds = list(range(2000, 2064))
    with mp.get_context("fork").Pool(min(16, os.cpu_count() or 1)) as pool:
        refs = pool.map(stress_parity.ref, seeds, chunksize=1)
    subs = [cohort_member(i) for i in seeds]
    imgs, masks, voxs = zip(*subs)
    shapes = [im.shape for im in imgs]
    handles = cohort.table_handles(ctx, shapes, voxs)
    out = ctx.pipeline_host(shapes, cohort.ALL_STAGES, image=np.concatenate([i.ravel() for i in imgs]),
                            mask=np.concatenate([m.ravel() for m in masks]), handles=handles, want_logbias=True)
    assert out["rc"] == 0
    off = 0
    for k, (i, it_ref, lb_ref) in enumerate(refs):
        V = imgs[k].size
        assert out["stats"]["n4_iters"][k][:4].tolist() == it_ref, (i, out["stats"]["n4_iters"][k][:4].tolist(), it_ref)
        lb = out["logbias"][off:off + V].reshape(shapes[k]).astype(np.float64)
        assert np.abs(np.expm1(lb - lb_ref)).max() < helpers.N4_BIAS_RTOL
        if k % 8 == 0:
            n4 = out["n4"][off:off + V].reshape(shapes[k])
            r = vdp_oracle.vdp(n4, masks[k].astype(np.float64), voxs[k])
            assert np.array_equal(out["defect"][off:off + V].reshape(shapes[k]), r["defectArray"])
            assert np.array_equal(out["defect_lb"][off:off + V].reshape(shapes[k]), r["defectArrayLB"])
            assert out["stats"]["mean"][k] == r["mean"] and out["stats"]["p99"][k] == r["p99"]
            ci_ref, shell_ref = ci_oracle.calculate_ci(r["defectArray"], voxs[k], generate(voxs[k], 50))
            sel = r["defectArray"] > 0
            assert np.array_equal(out["ci_shell"][off:off + V].reshape(shapes[k])[sel].astype(np.int32), shell_ref[sel])
        off += V


def test_resident_batch_in_parts_on_two_streams_changes_nothing(ctx, monkeypatch):
    """vb_pipeline_batch_dev splits large resident batches into parts that alternate between two
    streams (the tail of one part's persistent N4 kernel is filled by the next part and by the small
    kernels of the earlier ones): every output array identical to the unsplit run."""
    from vent_analysis_b200 import cohort
    subs = [cohort_member(i) for i in range(40, 54)]
    imgs, masks, voxs = zip(*subs)
    res = {}
    for parts in ("1", "2", "3", "4"):
        monkeypatch.setenv("VB_PARTS", parts)
        dc = cohort.DeviceCohort(imgs, masks, voxs, want_logbias=True)
        for _ in range(2):
            dc.step()
        dc.torch.cuda.synchronize()
        res[parts] = (dc.stats().tobytes(), dc.n4.cpu().numpy(), dc.defect.cpu().numpy(), dc.defect_lb.cpu().numpy(),
                      dc.defect_border.cpu().numpy(), dc.ci_shell.cpu().numpy())
    for parts in ("2", "3", "4"):
        assert res[parts][0] == res["1"][0]
        for a, b in zip(res[parts][1:], res["1"][1:]):
            assert np.array_equal(a, b)


def test_host_pipelines_in_many_chunks_over_four_streams_change_nothing(ctx, monkeypatch):
    """The host pipelines cut a batch into chunks whose N4 kernels rotate over four compute streams while
    the stages before / after N4 run on high-priority streams of their own: with chunks of three
    subjects (ten or more chunks in flight over all the streams) every output is identical to the
    one-chunk run, call after call, for the dense and for the compact entry point."""
    from vent_analysis_b200 import cohort
    subs = [cohort_member(i) for i in range(60, 90)]
    imgs, masks, voxs = zip(*subs)
    shapes = np.ascontiguousarray([im.shape for im in imgs], dtype=np.int32)
    handles = cohort.table_handles(ctx, shapes, voxs)
    img = np.concatenate([np.rint(np.clip(im, 0, 65535)).astype(np.float32).ravel() for im in imgs])
    msk = np.concatenate([m.ravel() for m in masks])
    p = ctx.default_n4_params(max_iter=(6, 4, 3))
    flags = cohort.ALL_STAGES | _lib.DO_MASK_BORDER
    keys = ("n4", "defect", "defect_lb", "defect_border", "mask_border", "ci_shell")
    monkeypatch.setenv("VB_CHUNK_PER", "100000")                       # one chunk
    ref = ctx.pipeline_host(shapes, flags, image=img, mask=msk, handles=handles, params=p)
    assert ref["rc"] == 0
    monkeypatch.setenv("VB_CHUNK_PER", "3")
    for rep in range(3):
        got = ctx.pipeline_host(shapes, flags, image=img, mask=msk, handles=handles, params=p)
        assert got["rc"] == 0 and got["stats"].tobytes() == ref["stats"].tobytes(), rep
        for k in keys:
            assert np.array_equal(got[k], ref[k]), (k, rep)
        pk = ctx.pipeline_host_packed(shapes, flags, image_f32=img, mask_u8=msk, handles=handles, params=p)
        assert pk["rc"] == 0
        un = _lib.unpack_outputs(pk, shapes)
        assert un["stats"].tobytes() == ref["stats"].tobytes(), rep
        for k in keys:
            assert np.array_equal(un[k], ref[k]), (k, rep)
