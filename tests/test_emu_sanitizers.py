"""Sanitizer pass (CPU): the product's host code + kernel sources, built for the host with
-fsanitize=address (tests/emu), run the degenerate / ragged batch and a golden VDP+CI case.
Catches heap misuse in the C-ABI layer (double frees, workspace overruns) that parity tests
cannot see.  Runs in a subprocess because libasan has to be preloaded into the interpreter."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from emu.build_emu import SMALL_BLOCKS

_SCRIPT = r'''
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import helpers
from vent_analysis_b200 import _lib
ctx = _lib.Context(0, lib=_lib.bind(sys.argv[2]))
helpers.check_edge_cases(ctx)
helpers.check_vdp_ci_against_golden(ctx, "small_b", True)
helpers.check_vdp_ci_against_golden(ctx, "ragged_c", False)
# a cluster launch is refused in this build (device memory on the instrumented heap): one CTA per subject runs instead
from vent_analysis_b200.phantom import phantom
img, mask = phantom((24, 20, 5), [1.5, 1.5, 10.0], 21)
ctx.set_n4_cluster(2)
out = ctx.pipeline_host([img.shape], _lib.DO_N4, image=img.ravel(), mask=mask.ravel(), params=ctx.default_n4_params(max_iter=(2, 1)))
assert out["rc"] == 0 and ctx.n4_cluster_size() == 1
ctx.set_n4_cluster(0)
ctx.close()
print("asan-ok")
'''


_TSAN_SCRIPT = r'''
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import numpy as np
from vent_analysis_b200 import _lib
from vent_analysis_b200.phantom import phantom
from vent_analysis_b200.sphere_table import generate
ctx = _lib.Context(0, lib=_lib.bind(sys.argv[2]))
shape, vox = (40, 36, 6), [1.5, 1.5, 10.0]
img, mask = phantom(shape, vox, 5)
tbl = generate(vox, 50)
h = ctx.ci_table(tbl, shape, key="tsan")
p = ctx.default_n4_params(max_iter=(4, 3, 2))      # not in debug mode: chain and next-iteration phases overlap
out = ctx.pipeline_host([shape, shape], _lib.DO_SNR | _lib.DO_N4 | _lib.DO_VDP | _lib.DO_CI | _lib.DO_MASK_BORDER,
                        image=np.concatenate([img.ravel(), img.ravel()]), mask=np.concatenate([mask.ravel(), mask.ravel()]),
                        handles=[h, h], params=p)
assert out["rc"] == 0
ctx.close()
print("tsan-ok")
'''


def test_kernels_are_race_free_under_tsan(tmp_path):
    """The emulation runs one OS thread per CUDA thread with std::barrier for __syncthreads /
    named barriers, so ThreadSanitizer checks the kernels' shared/global-memory accesses for
    missing barriers -- in particular the warp-team overlap of the N4 convergence chain with the
    speculative next-iteration phases."""
    libtsan = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libtsan) or not os.path.exists(libtsan):
        pytest.skip("libtsan not available")
    so = os.path.join(ROOT, "tests", "emu", "_build", "libventb200_emu_tsan.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    src = os.path.join(ROOT, "vent_analysis_b200", "csrc", "ventb200.cu")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-DVB_EMULATE", *SMALL_BLOCKS,
                           "-ffp-contract=off", "-fsanitize=thread", "-fno-omit-frame-pointer",
                           "-Wno-unknown-pragmas", "-x", "c++", src, "-o", so])
    script = tmp_path / "run.py"
    script.write_text(_TSAN_SCRIPT)
    env = dict(os.environ, LD_PRELOAD=libtsan, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=2 exitcode=0")
    r = subprocess.run([sys.executable, str(script), ROOT, so], env=env, capture_output=True, text=True, timeout=1500)
    assert "tsan-ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-6000:]


def test_host_layer_and_kernels_are_asan_clean(tmp_path):
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("libasan not available")
    so = os.path.join(ROOT, "tests", "emu", "_build", "libventb200_emu_asan.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    src = os.path.join(ROOT, "vent_analysis_b200", "csrc", "ventb200.cu")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-DVB_EMULATE", *SMALL_BLOCKS,
                           "-ffp-contract=off", "-fsanitize=address", "-fno-omit-frame-pointer",
                           "-Wno-unknown-pragmas", "-x", "c++", src, "-o", so])
    script = tmp_path / "run.py"
    script.write_text(_SCRIPT)
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0")
    r = subprocess.run([sys.executable, str(script), ROOT, so], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "asan-ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
