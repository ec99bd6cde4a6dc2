"""Sanitizer pass (CPU): the product's host code + kernel sources, built for the host with
-fsanitize=address (tests/emu), run the degenerate / ragged batch and a golden VDP+CI case.
Catches heap misuse in the C-ABI layer (double frees, workspace overruns) that parity tests
cannot see.  Runs in a subprocess because libasan has to be preloaded into the interpreter."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

_SCRIPT = r'''
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import helpers
from vent_analysis_b200 import _lib
ctx = _lib.Context(0, lib=_lib.bind(sys.argv[2]))
helpers.check_edge_cases(ctx)
helpers.check_vdp_ci_against_golden(ctx, "small_b", True)
helpers.check_vdp_ci_against_golden(ctx, "ragged_c", False)
ctx.close()
print("asan-ok")
'''


def test_host_layer_and_kernels_are_asan_clean(tmp_path):
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("libasan not available")
    so = os.path.join(ROOT, "tests", "emu", "_build", "libventb200_emu_asan.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    src = os.path.join(ROOT, "vent_analysis_b200", "csrc", "ventb200.cu")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-DVB_EMULATE",
                           "-ffp-contract=off", "-fsanitize=address", "-fno-omit-frame-pointer",
                           "-Wno-unknown-pragmas", "-x", "c++", src, "-o", so])
    script = tmp_path / "run.py"
    script.write_text(_SCRIPT)
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0")
    r = subprocess.run([sys.executable, str(script), ROOT, so], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "asan-ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
