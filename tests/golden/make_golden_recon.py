"""Golden fixture for the k-space reconstruction.  ``process_RAW`` starts by reading a
Siemens twix file through mapVBVD (absent, and no raw data is shipped), so the method cannot be
called; its arithmetic (Vent_Analysis.py:536-539) is executed FROM THE REFERENCE'S SOURCE TEXT --
the four lines are extracted by number and ``exec``-ed on a stand-in object that carries ``raw_K``.

    python tests/golden/make_golden_recon.py          # needs /root/reference
"""
import os
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Vent_Analysis.py"


def main():
    src = open(REF).read().splitlines()
    i0 = next(i for i, l in enumerate(src) if "self.raw_HPvent = np.zeros((self.raw_K.shape))" in l)
    block = textwrap.dedent("\n".join(src[i0:i0 + 4]))
    assert "np.transpose(self.raw_HPvent,(1,0,2))[:,::-1,:]" in block, block
    out = {}
    rng = np.random.default_rng(21)
    for name, shape in (("pow2", (16, 8, 3)), ("odd", (15, 12, 2)), ("mixed", (12, 7, 4))):
        K = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
        K[shape[0] // 2, shape[1] // 2, :] += 40.0                       # a DC-like peak, as in k-space
        me = types.SimpleNamespace(raw_K=K)
        exec(block, {"np": np, "self": me})
        out[f"K_{name}"] = K
        out[f"img_{name}"] = me.raw_HPvent
    np.savez_compressed(os.path.join(HERE, "recon_small.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
