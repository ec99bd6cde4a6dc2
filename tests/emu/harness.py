"""TEST INFRASTRUCTURE: a ventb200 ``Context`` backed by the host-emulated build of the
kernel sources (see build_emu.py).  Only the test-suite imports this."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if HERE not in sys.path:
    sys.path.insert(0, HERE)

_ctx = None


def emu_context():
    global _ctx
    if _ctx is None:
        import build_emu
        from vent_analysis_b200 import _lib
        so = os.environ.get("VB_EMU_LIB") or build_emu.build()     # VB_EMU_LIB: e.g. an -fsanitize=address build
        _ctx = _lib.Context(0, lib=_lib.bind(so))
    return _ctx
