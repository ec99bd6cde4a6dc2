"""The C ABI: the CUDA library builds for sm_100a, loads, and exports every symbol that
include/ventb200.h declares; ctypes mirrors have the C sizes; without a CUDA device the
library refuses to create a context (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT
from vent_analysis_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ventb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.EXPORTS) == syms
    assert lib.vb_version() >= 100


def test_struct_sizes_match_the_c_side(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include "ventb200.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                    'sizeof(vb_n4_params),sizeof(vb_subject_stats),sizeof(vb_buffers),sizeof(vb_host_io));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    a, b, c, d = (int(x) for x in subprocess.check_output([str(exe)]).split())
    assert d == ctypes.sizeof(_lib.HostIO)
    assert a == ctypes.sizeof(_lib.N4Params)
    assert b == ctypes.sizeof(_lib.SubjectStats) == _lib.STATS_DTYPE.itemsize
    assert c == ctypes.sizeof(_lib.Buffers)


def test_library_is_sm100a_sass_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert lib.vb_device_count() == 0
    with pytest.raises(_lib.VbError) as e:
        _lib.Context(0)
    assert "no CUDA device" in str(e.value)
    from vent_analysis_b200 import Vent_Analysis
    import numpy as np
    with pytest.raises(RuntimeError):
        Vent_Analysis(xenon_array=np.ones((8, 8, 2), np.float32), mask_array=np.ones((8, 8, 2)), vox=[1, 1, 1])
