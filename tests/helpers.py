"""Shared checkers: run the ventb200 C ABI (CUDA library on the GPU box, host-emulated build
of the same kernel sources in the CPU suite) and compare with the oracle / golden fixtures."""
import contextlib
import io

import numpy as np

from conftest import golden, unpack
from oracle import ci_oracle, n4_oracle, vdp_oracle
from vent_analysis_b200 import _lib
from vent_analysis_b200.phantom import phantom
from vent_analysis_b200.sphere_table import generate

N4_BIAS_RTOL = 1e-3      # north_star: N4 bias field within 1e-3 relative error
VDP_PP_TOL = 0.05        # north_star: VDP within +-0.05 percentage points


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def golden_inputs(name):
    g = golden(f"vdp_{name}.npz")
    shape = tuple(int(x) for x in g["shape"])
    vox = [float(x) for x in g["vox"]]
    if "img" in g.files:
        img, mask, n4 = g["img"], g["mask"], g["n4"]
    else:
        img, mask = phantom(shape, vox, int(g["seed"]))
        n4 = n4_oracle.n4(img, mask)["out"]
    return g, shape, vox, img, mask, n4


def check_vdp_ci_against_golden(ctx, name, with_ci=True):
    """VDP / LB / borders / SNR / CI from the SAME N4 volume the reference was given:
    masks, codes, counts and CI bit-exact (north_star: 'bit-exact given the same normalised
    volume'); SNR to float32 rounding of the reference's own float32 reduction."""
    g, shape, vox, img, mask, n4 = golden_inputs(name)
    flags = _lib.DO_SNR | _lib.DO_VDP | _lib.DO_MASK_BORDER | (_lib.DO_CI if with_ci else 0)
    handles = None
    if with_ci:
        tbl = generate(vox, 50)
        handles = [ctx.ci_table(tbl, shape, key=(tuple(vox), 50, tbl.shape[0]))]
    out = ctx.pipeline_host([shape], flags, image=img, mask=mask, n4=n4, handles=handles)
    assert out["rc"] == 0
    st = out["stats"][0]
    defect = out["defect"].reshape(shape)
    assert np.array_equal(defect, unpack(g["defect_bits"], shape))
    assert np.array_equal(out["defect_lb"].reshape(shape), g["defectLB"])
    assert np.array_equal(out["defect_border"].reshape(shape), unpack(g["defectBorder_bits"], shape))
    assert np.array_equal(out["mask_border"].reshape(shape), unpack(g["maskBorder_bits"], shape))
    nmask = np.sum(mask.astype(np.float64))
    assert 100 * np.float64(st["n_defect"]) / nmask == float(g["VDP"])
    assert 100 * np.int64(st["n_lb12"]) / nmask == float(g["VDP_lb"])
    assert np.int64(st["n_defect"]) * np.prod(np.divide(vox, 10)) / 1000 == float(g["DefectVolume"])
    assert np.isclose(st["snr"], float(g["SNR"]), rtol=2e-6)
    if with_ci:
        radii = ctx.ci_radii(handles[0])
        sel = defect > 0
        ci = radii[out["ci_shell"].reshape(shape)[sel]] * np.min(vox)
        assert np.array_equal(ci, g["CI_values"])                 # float64, bit exact
        assert (out["ci_shell"].reshape(shape)[~sel] == _lib.CI_OFF).all()
        assert radii[int(st["ci_t95"])] * np.min(vox) == float(g["CI"])
    return out


def check_ci_case(ctx, name):
    g = golden(f"ci_{name}.npz")
    d, vox, rmax = g["defect"], [float(x) for x in g["vox"]], int(g["rmax"])
    tbl = generate(vox, rmax)
    h = ctx.ci_table(tbl, d.shape, key=(tuple(vox), rmax, tbl.shape[0]))
    shell, stats, rc = ctx.ci_host(d, [d.shape], [h])
    assert rc == 0
    ci = np.zeros(d.shape)
    sel = d > 0
    ci[sel] = ctx.ci_radii(h)[shell.reshape(d.shape)[sel]] * min(vox)
    assert np.array_equal(ci, g["CIarray"])
    assert ctx.ci_radii(h)[int(stats["ci_t95"][0])] * min(vox) == vdp_oracle.ci_percentile(g["CIarray"], d)


def check_n4(ctx, img, mask, max_iter=(50, 50, 50, 50), require_same_iters=True):
    """N4 against the (unpinned) CPU restatement: bias field within 1e-3 relative
    (|exp(lb_gpu - lb_ref) - 1| on every voxel), corrected image likewise."""
    shape = img.shape
    p = ctx.default_n4_params(max_iter=max_iter)
    out = ctx.pipeline_host([shape], _lib.DO_N4, image=img, mask=mask, params=p, want_logbias=True)
    ref = n4_oracle.n4(img, mask, max_iter=max_iter)
    st = out["stats"][0]
    iters = [int(x) for x in st["n4_iters"][:len(max_iter)]]
    if require_same_iters:
        assert iters == [int(x) for x in ref["iters"]], (iters, ref["iters"])
    lb = out["logbias"].reshape(shape)
    rel = np.abs(np.expm1(lb.astype(np.float64) - ref["logbias"].astype(np.float64)))
    assert rel.max() < N4_BIAS_RTOL, rel.max()
    o = out["n4"].reshape(shape)
    relo = np.abs(o - ref["out"]) / np.maximum(np.abs(ref["out"]), 1e-6)
    assert relo.max() < N4_BIAS_RTOL, relo.max()
    assert int(st["n4_n_mask"]) == int((mask == 1).sum())
    return out, ref, float(rel.max())


def check_n4_cluster(ctx, cases, ks=(2, 4, 8), max_iter=(50, 50, 50, 50), sms=None):
    """Cluster mode of the N4 kernel (K CTAs per subject, vb_set_n4_cluster): every output and the
    iteration counts are BIT-IDENTICAL to one CTA per subject -- the cross-CTA sums are integer
    (fixed point), min / max are exact, and the float sums that differ only feed a bound with margins.
    `cases`: list of batches, each a list of (image, mask)."""
    p = ctx.default_n4_params(max_iter=max_iter)
    try:
        for subs in cases:
            shapes = [s[0].shape for s in subs]
            img = np.concatenate([s[0].ravel() for s in subs])
            msk = np.concatenate([s[1].ravel() for s in subs])
            res = {}
            for k in (1,) + tuple(ks):
                ctx.set_n4_cluster(k)
                out = ctx.pipeline_host(shapes, _lib.DO_N4, image=img, mask=msk, params=p, want_logbias=True)
                used = ctx.n4_cluster_size()
                if sms is not None:
                    want = k
                    while want > 1 and len(subs) * want > sms:
                        want //= 2
                    assert used == want, (k, used, want)
                res[k] = (used, out["n4"].copy(), out["logbias"].copy(), out["stats"].copy())
            assert res[1][0] == 1
            for k in ks:
                assert res[k][1].tobytes() == res[1][1].tobytes(), f"N4 output differs at K={k}"
                assert res[k][2].tobytes() == res[1][2].tobytes(), f"log bias differs at K={k}"
                for f in ("n4_iters", "n4_last_conv", "n4_n_mask"):
                    assert res[k][3][f].tobytes() == res[1][3][f].tobytes(), (f, k)
    finally:
        ctx.set_n4_cluster(0)


def check_edge_cases(ctx):
    """Degenerate and ragged inputs through the whole path in ONE batch (odd voxel offsets, a
    single-slice volume, an empty mask, a one-voxel mask, a mask with a foreign label):
    N4 within tolerance of the restatement, everything downstream bit-exact given our N4."""
    rng = np.random.default_rng(3)
    cases = []
    img, mask = phantom((33, 29, 5), [1.5, 1.5, 10.0], 41)          # odd sizes -> unaligned offsets behind it
    cases.append((img, mask))
    img, mask = phantom((36, 40, 1), [2.0, 2.0, 11.5], 42)          # single slice
    cases.append((img, mask))
    img = (100 + 10 * rng.standard_normal((21, 23, 3))).astype(np.float32)
    cases.append((img, np.zeros((21, 23, 3), np.uint8)))            # empty mask: N4 is the identity
    one = np.zeros((21, 23, 3), np.uint8); one[10, 11, 1] = 1
    cases.append((img, one))                                         # one mask voxel: no fit possible
    img, mask = phantom((40, 36, 6), [1.5, 1.5, 10.0], 43)
    m2 = mask.copy(); m2[:6] = np.where(m2[:6] > 0, 2, 0)            # label 2 is outside N4's mask (MaskLabel = 1)
    cases.append((img, m2))
    shapes = [c[0].shape for c in cases]
    p = ctx.default_n4_params(max_iter=(6, 4))
    out = ctx.pipeline_host(shapes, _lib.DO_N4, image=np.concatenate([c[0].ravel() for c in cases]),
                            mask=np.concatenate([c[1].ravel() for c in cases]), params=p, want_logbias=True)
    off = 0
    for (img, mask), shape in zip(cases, shapes):
        V = img.size
        n4 = out["n4"][off:off + V].reshape(shape)
        ref = n4_oracle.n4(img, mask, max_iter=(6, 4))
        if (mask == 1).sum() < 2:
            assert np.array_equal(n4, img)                           # zero lattice: out = in / exp(0)
        else:
            lb = out["logbias"][off:off + V].reshape(shape).astype(np.float64)
            assert np.abs(np.expm1(lb - ref["logbias"])).max() < N4_BIAS_RTOL
        off += V
    # VDP / borders on the first two cases (non-empty 0/1 masks), given our N4
    off = 0
    for (img, mask), shape in list(zip(cases, shapes))[:2]:
        V = img.size
        n4 = out["n4"][off:off + V].reshape(shape)
        o = ctx.pipeline_host([shape], _lib.DO_VDP | _lib.DO_MASK_BORDER, mask=mask, n4=n4)
        r = vdp_oracle.vdp(n4, mask.astype(np.float64), [1.5, 1.5, 10.0])
        assert np.array_equal(o["defect"].reshape(shape), r["defectArray"])
        assert np.array_equal(o["defect_lb"].reshape(shape), r["defectArrayLB"])
        assert np.array_equal(o["defect_border"].reshape(shape).astype(bool), r["defectBorder"])
        assert np.array_equal(o["mask_border"].reshape(shape), vdp_oracle.border(mask))
        off += V


def _vdp_reference_given_stats(n4, mask, mean, p99, thresh=0.6):
    """numpy restatement of Vent_Analysis.py:246-256 for a GIVEN float32 mean / p99 (exact
    float32 divisions and comparisons), median and border from the oracle."""
    n4 = np.asarray(n4, dtype=np.float32)
    with np.errstate(all="ignore"):
        d0 = (np.divide(n4, np.float32(mean)) < thresh) * (mask != 0)
        x = np.divide(n4, np.float32(p99))
    e = vdp_oracle.LB_EDGES
    lb = ((x <= e[0]) * 1 + (x > e[0]) * (x <= e[1]) * 2 + (x > e[1]) * (x <= e[2]) * 3
          + (x > e[2]) * (x <= e[3]) * 4 + (x > e[3]) * (x <= e[4]) * 5 + (x > e[4]) * 6) * (mask != 0)
    defect = vdp_oracle.median3x3_binary(d0)
    return defect.astype(np.uint8), lb.astype(np.uint8), vdp_oracle.border(defect).astype(np.uint8)


def _near_threshold_volume(shape, seed, odd_in_mask=()):
    """A volume whose in-mask values sit within a few ulps of the quotient thresholds
    (0.6 * mean and the linear-binning edges * p99), with zero / negative / denormal voxels in
    the mask; `odd_in_mask` (NaN / inf) makes the mean or p99 degenerate."""
    rng = np.random.default_rng(seed)
    n4 = (1000.0 * (1 + 0.2 * rng.standard_normal(shape))).astype(np.float32)
    mask = (rng.random(shape) < 0.6).astype(np.uint8)
    sel = np.flatnonzero(mask.ravel())
    pick = rng.permutation(sel)
    edge, odd = pick[:len(sel) // 2], pick[len(sel) // 2:len(sel) // 2 + 7 + len(odd_in_mask)]
    cs = rng.choice(np.array([0.6, 0.16, 0.34, 0.52, 0.7, 0.88]), size=len(edge))
    flat = n4.ravel().astype(np.float64)
    flat[odd[:7]] = np.array([0.0, -0.0, -5.0, 1e-30, 7.0, -1e-40, 1e-45])
    for _ in range(60):                                   # fixed point: edge voxels relative to the final mean / p99
        cur = flat.astype(np.float32)
        vals = np.sort(cur[sel])
        mean = np.mean(vals)                              # the reference's anchor (V_A.py:245-246)
        p99 = vals[int(len(vals) * .99)]
        anchor = np.where(cs == 0.6, np.float64(mean), np.float64(p99))
        flat[edge] = anchor * cs
    out = flat.astype(np.float32)
    bump = rng.integers(-3, 4, size=len(edge))
    out[edge] = (out[edge].view(np.int32) + bump.astype(np.int32)).view(np.float32)
    if len(odd_in_mask):
        out[odd[7:]] = np.array(odd_in_mask, np.float32)
    out[np.flatnonzero(mask.ravel() == 0)[:4]] = np.array([np.nan, np.inf, -np.inf, 3e38], np.float32)
    return out.reshape(shape), mask


def _same_f32(a, b):
    return a == b or (np.isnan(a) and np.isnan(b))


def check_vdp_threshold_edges(ctx):
    """Threshold / linear-binning / median / border kernels on values placed a few ulps around
    every quotient threshold (word and byte versions, mixed in one batch) against the restatement
    of the reference (oracle/vdp_oracle.vdp): the anchors are the reference's own -- numpy's float32
    pairwise mean of the SORTED mask values and the exact p99 -- so every output is bit exact."""
    shapes = [(32, 20, 6), (35, 8, 3), (64, 12, 7), (16, 16, 14), (24, 8, 2), (20, 12, 1), (33, 9, 5), (21, 10, 6)]
    odd = [(), (), (), (), (np.inf,), (np.nan, -np.inf), (), (np.inf, np.nan)]        # last two: byte version (unaligned)
    vols = [_near_threshold_volume(s, 100 + i, o) for i, (s, o) in enumerate(zip(shapes, odd))]
    n4 = np.concatenate([v[0].ravel() for v in vols])
    mask = np.concatenate([v[1].ravel() for v in vols])
    out = ctx.pipeline_host(shapes, _lib.DO_VDP, mask=mask, n4=n4)
    off = 0
    for (x, m), shape, st in zip(vols, shapes, out["stats"]):
        V = x.size
        with np.errstate(all="ignore"):
            r = vdp_oracle.vdp(x, m.astype(np.float64), [1.5, 1.5, 10.0])
        assert int(st["n_mask"]) == int((m > 0).sum())
        assert _same_f32(st["p99"], r["p99"]), (shape, st["p99"], r["p99"])
        assert _same_f32(st["mean"], r["mean"]), (shape, st["mean"], r["mean"])
        assert np.array_equal(out["defect"][off:off + V].reshape(shape), r["defectArray"]), shape
        assert np.array_equal(out["defect_lb"][off:off + V].reshape(shape), r["defectArrayLB"]), shape
        assert np.array_equal(out["defect_border"][off:off + V].reshape(shape).astype(bool), r["defectBorder"]), shape
        assert 100 * np.float64(st["n_defect"]) / np.sum(m.astype(np.float64)) == r["VDP"]
        assert 100 * np.int64(st["n_lb12"]) / np.sum(m.astype(np.float64)) == r["VDP_lb"]
        off += V


def check_mean_bit_exact(ctx, n_phantoms=8, shape=(48, 40, 6), big=None):
    """stats.mean == np.mean(sorted(N4[mask>0])) (Vent_Analysis.py:245-246: numpy's float32 PAIRWISE
    sum of the sorted list / M) and stats.p99 == sorted[int(M*.99)], bit for bit: phantoms, mask
    sizes around every corner of numpy's recursion (M < 8, M <= 128, just above, odd remainders),
    duplicate-heavy and wide-range data, and the global-memory variant of the sort."""
    import os
    rng = np.random.default_rng(77)
    vols = []
    for i in range(n_phantoms):
        img, mask = phantom(shape, [2.0, 2.0, 11.5], 900 + i)
        vols.append((img, mask))
    for M in (1, 5, 7, 8, 9, 15, 16, 127, 128, 129, 136, 137, 255, 256, 257, 1000, 1029):
        x = (500 * (1 + 0.3 * rng.standard_normal((40, 30, 2)))).astype(np.float32)
        m = np.zeros(x.size, np.uint8)
        m[rng.permutation(x.size)[:M]] = 1
        vols.append((x, m.reshape(x.shape)))
    x = rng.choice(np.array([3.0, 3.0000002, 700.0, 1e-3, 2.5e7], np.float32), size=(30, 24, 3))      # few distinct values
    vols.append((x, (rng.random(x.shape) < 0.8).astype(np.uint8)))
    x = (rng.standard_normal((30, 24, 3)) * 10.0 ** rng.integers(-6, 6, size=(30, 24, 3))).astype(np.float32)   # signed, wide
    vols.append((x, (rng.random(x.shape) < 0.7).astype(np.uint8)))
    if big is not None:
        vols.append(big)
    shapes = [v[0].shape for v in vols]
    for cap in (None, "64"):
        if cap is None:
            os.environ.pop("VB_VDP_KEY_CAP", None)
        else:
            os.environ["VB_VDP_KEY_CAP"] = cap
        try:
            out = ctx.pipeline_host(shapes, _lib.DO_VDP, mask=np.concatenate([v[1].ravel() for v in vols]),
                                    n4=np.concatenate([v[0].ravel() for v in vols]))
        finally:
            os.environ.pop("VB_VDP_KEY_CAP", None)
        for (x, m), st in zip(vols, out["stats"]):
            vals = sorted(x[m > 0])                                   # the reference's expression, literally
            assert int(st["n_mask"]) == len(vals)
            assert st["mean"] == np.mean(vals), (cap, x.shape, len(vals), st["mean"], np.mean(vals))
            assert st["p99"] == vals[int(len(vals) * .99)], (cap, x.shape, len(vals))


def check_p99_select(ctx):
    """Exact order statistic on clustered / duplicate-heavy / signed data, keys in shared memory
    and (VB_VDP_KEY_CAP) re-streamed from global memory."""
    import os
    rng = np.random.default_rng(9)
    shape = (24, 16, 5)
    cases = [np.full(shape, 3.25, np.float32),
             rng.choice(np.array([1.0, 1.0000001, 2.0, -3.0], np.float32), size=shape),
             (1000 + rng.standard_normal(shape)).astype(np.float32),
             (rng.standard_normal(shape) * 10.0 ** rng.integers(-20, 20, size=shape)).astype(np.float32),
             np.abs(rng.standard_normal(shape)).astype(np.float32) * np.float32(1e-40)]
    masks = [(rng.random(shape) < p).astype(np.uint8) for p in (1.0, 0.5, 0.3, 0.9, 0.7)]
    masks[2][0, 0, 0] = 1
    for cap in (None, "64"):
        if cap is None:
            os.environ.pop("VB_VDP_KEY_CAP", None)
        else:
            os.environ["VB_VDP_KEY_CAP"] = cap
        try:
            out = ctx.pipeline_host([shape] * len(cases), _lib.DO_VDP, mask=np.concatenate([m.ravel() for m in masks]),
                                    n4=np.concatenate([c.ravel() for c in cases]))
        finally:
            os.environ.pop("VB_VDP_KEY_CAP", None)
        for x, m, st in zip(cases, masks, out["stats"]):
            vals = np.sort(x[m > 0])
            assert int(st["n_mask"]) == len(vals)
            assert st["p99"] == vals[int(len(vals) * .99)], (cap, st["p99"], vals[int(len(vals) * .99)])


def check_pack6(ctx):
    """vb_pack6_batch_host (Vent_Analysis.build4DdataArray, V_A.py:292-313) against the numpy
    statement of the reference: ragged batch, one subject without CI values, proton present /
    absent, float32 bit-exact (CIarray = float32(float64 radius * min(vox)))."""
    rng = np.random.default_rng(4)
    vox = [1.5, 1.5, 10.0]
    tbl = generate(vox, 50)
    shapes = [(12, 10, 3), (9, 7, 5), (16, 8, 2)]
    hs = [ctx.ci_table(tbl, s, key=(tuple(vox), 50, tbl.shape[0])) for s in shapes]
    radii = ctx.ci_radii(hs[0])
    vols = []
    for s in shapes:
        img = (1000 * rng.random(s)).astype(np.float32)
        pro = (500 * rng.random(s)).astype(np.float32)
        mask = (rng.random(s) < 0.5).astype(np.uint8)
        n4 = (img * (1 + 0.1 * rng.random(s))).astype(np.float32)
        defect = ((rng.random(s) < 0.2) & (mask > 0)).astype(np.uint8)
        shell = np.where(defect > 0, rng.integers(0, len(radii), size=s), _lib.CI_OFF).astype(np.uint16)
        vols.append((pro, img, mask, n4, defect, shell))
    cat = lambda k: np.concatenate([v[k].ravel() for v in vols])
    for with_proton in (True, False):
        out = ctx.pack6_host(shapes, [min(vox)] * 3, proton=cat(0) if with_proton else None, image=cat(1), mask=cat(2),
                             n4=cat(3), defect=cat(4), ci_shell=cat(5), handles=hs)
        off = 0
        for (pro, img, mask, n4, defect, shell), s in zip(vols, shapes):
            V = img.size
            ref = np.zeros(s + (6,), dtype=np.float32)                # V_A.py:294
            if with_proton:
                ref[..., 0] = pro
            ref[..., 1] = img
            ref[..., 2] = mask.astype(np.float64)
            ref[..., 3] = n4
            ref[..., 4] = defect.astype(np.float64)
            ci = np.zeros(s)                                          # CI.py:139-142: float64 radius * min(vox)
            sel = defect > 0
            ci[sel] = radii[shell[sel]] * min(vox)
            ref[..., 5] = ci
            assert np.array_equal(out[6 * off:6 * (off + V)].reshape(s + (6,)), ref), (s, with_proton)
            off += V
    # channels whose input is missing stay zero
    out = ctx.pack6_host(shapes, [min(vox)] * 3, image=cat(1), mask=cat(2))
    assert not out.reshape(-1, 6)[:, [0, 3, 4, 5]].any() and np.array_equal(out.reshape(-1, 6)[:, 1], cat(1))


def check_pack6_golden(ctx):
    """The packing against the UNMODIFIED reference's build4DdataArray (tests/golden/pack6_small_b.npz,
    made by tests/golden/make_golden_pack6.py): VDP + CI through our kernels on the fixture's
    inputs, packed on the device, SHA-256 of the float32 [48, 40, 6, 6] array identical."""
    import hashlib
    gp = golden("pack6_small_b.npz")
    g, shape, vox, img, mask, n4 = golden_inputs("small_b")
    out = check_vdp_ci_against_golden(ctx, "small_b", True)
    tbl = generate(vox, 50)
    h = ctx.ci_table(tbl, shape, key=(tuple(vox), 50, tbl.shape[0]))
    proton = (500.0 * np.random.default_rng(int(gp["proton_seed"])).random(shape)).astype(np.float32)
    packed = ctx.pack6_host([shape], [min(vox)], proton=proton, image=img, mask=mask, n4=n4, defect=out["defect"],
                            ci_shell=out["ci_shell"], handles=[h])
    assert tuple(int(x) for x in gp["shape"]) == shape + (6,) and str(gp["dtype"]) == "float32"
    assert np.allclose(packed.reshape(-1, 6).astype(np.float64).sum(axis=0), gp["channel_sums"], rtol=0, atol=0)
    assert hashlib.sha256(packed.tobytes()).hexdigest() == str(gp["sha"])


def check_packed_io(ctx, shapes=((33, 29, 5), (24, 20, 4), (36, 40, 1), (40, 32, 6)), max_iter=(4, 3)):
    """vb_pipeline_batch_host_packed (uint16 image + bit-packed mask in; N4HPvent, code bytes and the
    sparse CI list out) against vb_pipeline_batch_host on the same ragged batch: every array of the
    class surface identical after unpacking; float32-image / byte-mask inputs likewise; a CI
    capacity that is too small is an argument error."""
    import pytest
    vox = [1.5, 1.5, 10.0]
    subs = [phantom(s, vox, 60 + i) for i, s in enumerate(shapes)]
    u16 = [np.rint(np.clip(im, 0, 65535)).astype(np.uint16) for im, _ in subs]      # DICOM pixel data
    masks = [m for _, m in subs]
    img = np.concatenate([u.astype(np.float32).ravel() for u in u16])                # V_A.py:322 cast
    msk = np.concatenate([m.ravel() for m in masks])
    tbl = generate(vox, 50)
    handles = [ctx.ci_table(tbl, s, key=(tuple(vox), 50, tbl.shape[0])) for s in shapes]
    p = ctx.default_n4_params(max_iter=max_iter)
    flags = _lib.DO_SNR | _lib.DO_N4 | _lib.DO_VDP | _lib.DO_CI | _lib.DO_MASK_BORDER
    ref = ctx.pipeline_host(shapes, flags, image=img, mask=msk, handles=handles, params=p, want_logbias=True)
    assert ref["rc"] == 0
    keys = ("n4", "logbias", "defect", "defect_lb", "defect_border", "mask_border", "ci_shell")
    for kw in (dict(image_u16=np.concatenate([u.ravel() for u in u16]), mask_bits=_lib.pack_mask_bits(masks)),
               dict(image_f32=img, mask_u8=msk)):
        pk = ctx.pipeline_host_packed(shapes, flags, handles=handles, params=p, want_logbias=True, **kw)
        assert pk["rc"] == 0
        got = _lib.unpack_outputs(pk, shapes)
        for k in keys:
            assert np.array_equal(got[k], ref[k]), (k, list(kw))
        assert got["stats"].tobytes() == ref["stats"].tobytes()
        D = ref["stats"]["ci_n"].astype(np.int64)
        assert np.array_equal(pk["ci_offset"], np.concatenate([[0], np.cumsum(D)]))
    # scalars only: no arrays wanted
    pk = ctx.pipeline_host_packed(shapes, flags & ~_lib.DO_MASK_BORDER, image_f32=img, mask_u8=msk, handles=handles, params=p,
                                  want_n4=False, want_codes=False, want_ci_map=False)
    assert set(pk) == {"stats", "rc"} and np.array_equal(pk["stats"]["ci_t95"], ref["stats"]["ci_t95"])
    assert np.array_equal(pk["stats"]["n_defect"], ref["stats"]["n_defect"])
    small = dict(ci_index=np.empty(3, np.uint32), ci_value=np.empty(3, np.uint16))
    with pytest.raises(_lib.VbError) as e:
        ctx.pipeline_host_packed(shapes, flags, image_f32=img, mask_u8=msk, handles=handles, params=p, buffers=small)
    assert e.value.code == _lib.VB_ERR_ARG


def check_overlays(ctx):
    """Overlay compositions through the C ABI against arrays produced by the reference's
    own exportDICOM / cropToData / screenShot / colorBinary (tests/golden/overlay_small_b.npz), bit
    exact; plus a ragged batch against the oracle, the float64 (integer image) mode, a flat image,
    and the two IndexError cases of the reference."""
    import pytest
    from oracle import overlay_oracle
    g = golden("overlay_small_b.npz")
    v, shape, vox, img, mask, n4 = golden_inputs("small_b")
    out = check_vdp_ci_against_golden(ctx, "small_b", True)
    defect, mborder, shell = out["defect"].reshape(shape), out["mask_border"].reshape(shape), out["ci_shell"].reshape(shape)
    tbl = generate(vox, 50)
    h = ctx.ci_table(tbl, shape, key=(tuple(vox), 50, tbl.shape[0]))
    proton = (500.0 * np.random.default_rng(int(g["proton_seed"])).random(shape)).astype(np.float32)
    assert np.array_equal(ctx.overlay_rgb_host([shape], n4, defect).reshape(shape + (3,)), g["export_rgb"])
    assert np.array_equal(ctx.overlay_rgb_host([shape], n4, defect, gui=True).reshape(shape + (3,)), g["gui_defect"])
    assert np.array_equal(ctx.overlay_rgb_host([shape], img, mborder, gui=True).reshape(shape + (3,)), g["gui_raw_border"])
    assert [int(x) for x in ctx.crop_boxes_host([shape], mask, border=5)[0]] == [int(x) for x in g["crop_box"]]
    imgs, boxes = ctx.montage_host([shape], img, n4, mask, mborder, defect, g["parula"], proton=proton, ci_shell=shell,
                                   handles=[h], vox_min=[min(vox)])
    assert imgs[0].shape == g["screenshot"].shape and np.array_equal(imgs[0], g["screenshot"])
    # ragged batch, integer-valued image (float64 normalisation), against the oracle
    rng = np.random.default_rng(8)
    shapes = [(20, 17, 4), (33, 12, 3), (16, 16, 5)]
    vols = []
    for i, s_ in enumerate(shapes):
        im, mk = phantom(s_, vox, 70 + i)
        u16 = np.rint(im).astype(np.uint16)
        x4 = (im * (1 + 0.1 * rng.random(s_))).astype(np.float32)
        df = ((rng.random(s_) < 0.3) & (mk > 0)).astype(np.uint8)
        vols.append((u16, mk, x4, df, vdp_oracle.border(mk).astype(np.uint8), (300 * rng.random(s_)).astype(np.float32)))
    cat = lambda k, dt=None: np.concatenate([np.asarray(v_[k], dtype=dt).ravel() for v_ in vols])
    rgb = ctx.overlay_rgb_host(shapes, cat(2), cat(3))
    rgb_u16 = ctx.overlay_rgb_host(shapes, cat(0), cat(4), gui=True)
    boxes = ctx.crop_boxes_host(shapes, cat(1), border=2, border_slices=True)
    imgs, mboxes = ctx.montage_host(shapes, cat(0), cat(2), cat(1), cat(4), cat(3), g["parula"], proton=cat(5))
    off = 0
    for i, ((u16, mk, x4, df, bd, pr), s_) in enumerate(zip(vols, shapes)):
        V = u16.size
        assert np.array_equal(rgb[3 * off:3 * (off + V)].reshape(s_ + (3,)), overlay_oracle.export_rgb(x4, df)), s_
        assert np.array_equal(rgb_u16[3 * off:3 * (off + V)].reshape(s_ + (3,)), overlay_oracle.color_binary(u16, bd)), s_
        assert [int(x) for x in boxes[i]] == list(overlay_oracle.crop_box(mk, border=2, border_slices=True))
        ref = overlay_oracle.screenshot_image(u16, pr, x4, mk.astype(np.float64), bd, df, np.zeros(s_), g["parula"])
        assert np.array_equal(imgs[i], ref), s_
        off += V
    # a flat image is returned as it is by normalize(): uint8 wrap of x * 255
    flat = np.full((8, 8, 2), 3.0, np.float32)
    z = np.zeros((8, 8, 2), np.uint8)
    assert np.array_equal(ctx.overlay_rgb_host([(8, 8, 2)], flat, z).reshape(8, 8, 2, 3), overlay_oracle.export_rgb(flat, z))
    # reference IndexErrors: no index > 0 with signal; a CI value >= 40 mm
    with pytest.raises(IndexError):
        ctx.crop_boxes_host([(8, 8, 2)], z)
    big = np.zeros(shape, np.uint16) + _lib.CI_OFF
    big[defect > 0] = len(ctx.ci_radii(h)) - 1                        # 49.99 voxels * 1.5 mm
    with pytest.raises(IndexError):
        ctx.montage_host([shape], img, n4, mask, mborder, defect, g["parula"], proton=proton, ci_shell=big, handles=[h],
                         vox_min=[min(vox)])


RECON_RTOL = 1e-5        # relative to the largest magnitude of the image (single-precision transforms)


def check_recon(ctx):
    """vb_recon_batch_host (process_RAW, Vent_Analysis.py:536-540) against images produced by the
    reference's own source lines (tests/golden/recon_small.npz) and, on a ragged batch with
    power-of-two, odd and DICOM-like (52 x 64) matrices, against the oracle: within 1e-5 of the
    largest magnitude."""
    from oracle import recon_oracle
    g = golden("recon_small.npz")
    names = ("pow2", "odd", "mixed")
    imgs = ctx.recon_host([g[f"K_{n}"] for n in names])
    for n, im in zip(names, imgs):
        ref = g[f"img_{n}"]
        assert im.shape == ref.shape
        assert np.abs(im - ref).max() <= RECON_RTOL * np.abs(ref).max(), n
    rng = np.random.default_rng(5)
    ks = []
    for shape in ((64, 64, 5), (52, 64, 18), (33, 20, 1), (128, 128, 14)):
        K = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
        K[shape[0] // 2, shape[1] // 2] += 300.0
        ks.append(K)
    for K, im in zip(ks, ctx.recon_host(ks)):
        ref = recon_oracle.reconstruct(K)
        assert np.abs(im - ref).max() <= RECON_RTOL * np.abs(ref).max(), K.shape
    # linearity and Parseval as size-independent properties
    a, b = ks[3], np.roll(ks[3], 3, axis=0)
    ia, ib, iab = ctx.recon_host([a, b, (a + 2 * b).astype(np.complex64)])
    assert np.abs(iab - (ia + 2 * ib)).max() <= 4 * RECON_RTOL * np.abs(iab).max()
    assert np.isclose(np.sum(np.abs(ia) ** 2), a.shape[0] * a.shape[1] * np.sum(np.abs(a) ** 2), rtol=1e-4)
