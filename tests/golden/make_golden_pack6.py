"""Golden fixture for the channel packing: runs the UNMODIFIED reference's
``Vent_Analysis.build4DdataArray`` (Vent_Analysis.py:292-313) in the build container on the inputs
of ``vdp_small_b.npz`` (N4 pinned to the stored volume, CI by the reference's own CI.py) with a
seeded proton volume, and stores the SHA-256 of the float32 ``[48, 40, 6, 6]`` result.

    python tests/golden/make_golden_pack6.py          # needs /root/reference
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, load_reference, run_reference, sha  # noqa: E402


def proton_for(shape, seed=99):
    return (500.0 * np.random.default_rng(seed).random(shape)).astype(np.float32)


def main():
    VA, _ = load_reference()
    g = np.load(os.path.join(HERE, "vdp_small_b.npz"), allow_pickle=True)
    shape = tuple(int(x) for x in g["shape"])
    vox = [float(x) for x in g["vox"]]
    v = run_reference(VA, g["img"], g["mask"].astype(np.float64), vox, g["n4"], True, REF)
    v.proton = proton_for(shape)
    arr = v.build4DdataArray()
    assert arr.dtype == np.float32 and arr.shape == shape + (6,)
    np.savez_compressed(os.path.join(HERE, "pack6_small_b.npz"), sha=sha(arr), shape=np.asarray(arr.shape),
                        dtype=str(arr.dtype), proton_seed=99, channel_sums=arr.reshape(-1, 6).astype(np.float64).sum(axis=0))
    print("pack6_small_b", arr.shape, sha(arr), arr.reshape(-1, 6).astype(np.float64).sum(axis=0))


if __name__ == "__main__":
    main()
