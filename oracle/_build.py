"""Builds / loads ``oracle/_build/liboracle.so`` (the C restatements).  TEST
INFRASTRUCTURE; the product never imports this."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRC = [os.path.join(_HERE, f) for f in ("ci_oracle.c", "n4_oracle.c")]
_lib = None


def build(force=False):
    stale = force or not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in _SRC)
    if stale:
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib
