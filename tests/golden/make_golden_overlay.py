"""Golden fixture for the overlay compositions: arrays produced by the UNMODIFIED
reference methods ``exportDICOM`` (Vent_Analysis.py:380-391), ``cropToData`` (:430-456), ``screenShot``
(:458-494) and the GUI helper arithmetic of ``colorBinary`` (:628-634) on the ``small_b`` fixture.

    python tests/golden/make_golden_overlay.py          # needs /root/reference

The reference methods are run as they are; what is absent from the container is stubbed AROUND them:
  * ``pydicom`` (module stub; ``exportDICOM`` is given a plain object as ``ds`` and the bytes it
    stores in ``ds.PixelData`` are the fixture -- the ``forPACS=False`` branch);
  * ``skimage.util.montage`` (absent): the grid placement ``oracle.overlay_oracle.montage`` (k-th image
    to cell (k // ncols, k % ncols), no padding) -- the only non-reference arithmetic on this path;
  * ``np.load('C:\\PIRL\\data\\parula.np.npy')`` is redirected to the reference's own ``parula.npy``;
  * ``ImageFont.truetype`` (no arial.ttf here) -> PIL's default font; the array is captured where the
    reference hands it to ``Image.fromarray`` (before any text is drawn).
``colorBinary`` lives inside ``if __name__ == '__main__'`` and cannot be imported: its six lines are
executed from the reference SOURCE TEXT (extracted by line range and ``exec``-ed), not retyped.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = "/root/reference"


def main():
    import make_golden as mg
    from oracle import overlay_oracle
    VA, CI = mg.load_reference()
    sys.modules['skimage.util'].montage = lambda imgs, grid_shape, padding_width=0, fill=0: overlay_oracle.montage(list(imgs), grid_shape)
    sys.modules['pydicom'].uid = types.SimpleNamespace(generate_uid=lambda: "1.2.3")

    g = np.load(os.path.join(HERE, "vdp_small_b.npz"))
    shape = tuple(int(x) for x in g["shape"])
    vox = [float(x) for x in g["vox"]]
    img, mask, n4 = g["img"], g["mask"], g["n4"]
    v = mg.run_reference(VA, img, mask.astype(np.float64), vox, n4, True, REF)
    proton = (500.0 * np.random.default_rng(11).random(shape)).astype(np.float32)
    v.proton = proton
    out = {"proton_seed": 11}

    # exportDICOM, forPACS=False: ds.PixelData = transpose(RGB, (2,0,1,3)).tobytes()
    ds = types.SimpleNamespace()
    ds.save_as = lambda path: None
    v.metadata['PatientName'] = 'golden'
    mg.quiet(v.exportDICOM, ds, "/tmp", "", False)
    rgb = np.frombuffer(ds.PixelData, dtype=np.uint8).reshape(shape[2], shape[0], shape[1], 3).transpose(1, 2, 0, 3)
    out["export_rgb"] = np.ascontiguousarray(rgb)

    # cropToData(mask, border=5)
    _, rr, cc, ss = v.cropToData(v.mask, border=5)
    out["crop_box"] = np.array([rr[0], rr[-1] + 1, cc[0], cc[-1] + 1, ss[0], ss[-1] + 1], dtype=np.int32)

    # screenShot: capture the array handed to PIL
    from PIL import Image, ImageFont
    captured = {}
    real_fromarray, real_truetype, real_load = Image.fromarray, ImageFont.truetype, np.load
    def fromarray(a, *k, **kw):
        captured["image"] = np.array(a)
        return real_fromarray(a, *k, **kw)
    Image.fromarray = fromarray
    default_font = ImageFont.load_default()
    ImageFont.truetype = lambda *a, **k: default_font
    np.load = lambda p, *a, **k: real_load(os.path.join(REF, "parula.npy")) if "parula" in str(p) else real_load(p, *a, **k)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mg.quiet(v.screenShot, "/tmp/vb_golden_screenshot.png")
    finally:
        Image.fromarray, ImageFont.truetype, np.load = real_fromarray, real_truetype, real_load
    out["screenshot"] = captured["image"]
    out["parula"] = real_load(os.path.join(REF, "parula.npy"))       # the colour table the reference loaded (64 x 3 float64)

    # colorBinary (GUI helper, Vent_Analysis.py:622-634), executed from the reference's source text
    src = open(os.path.join(REF, "Vent_Analysis.py")).read().splitlines()
    start = next(i for i, l in enumerate(src) if l.strip().startswith("def normalize(x):") and i > 600)
    end = next(i for i, l in enumerate(src) if l.strip().startswith("def colorparula"))
    import textwrap
    ns = {"np": np}
    exec(textwrap.dedent("\n".join(src[start:end])), ns)
    # the GUI applies it to 1 x slices montages of abs(volume) (:642-643, :734-748)
    mont = lambda A: overlay_oracle.montage([abs(A[:, :, k]) for k in range(A.shape[2])], (1, A.shape[2]))
    cb = ns["colorBinary"](mont(v.N4HPvent), mont(v.defectArray)).astype(np.uint8)           # arrayToImage: astype(np.uint8)
    out["gui_defect"] = np.stack([cb[:, k * shape[1]:(k + 1) * shape[1], :] for k in range(shape[2])], axis=2)
    cb = ns["colorBinary"](mont(v.HPvent), mont(v.mask_border)).astype(np.uint8)
    out["gui_raw_border"] = np.stack([cb[:, k * shape[1]:(k + 1) * shape[1], :] for k in range(shape[2])], axis=2)
    np.savez_compressed(os.path.join(HERE, "overlay_small_b.npz"), **out)
    print({k: (getattr(val, "shape", None), getattr(val, "dtype", None)) for k, val in out.items()})


if __name__ == "__main__":
    main()
