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
