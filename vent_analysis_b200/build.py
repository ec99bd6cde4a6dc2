"""Builds ``vent_analysis_b200/lib/libventb200.so`` (hand-written CUDA, sm_100a only) with
nvcc.  In-tree so that the built library travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libventb200.so")
SOURCES = ["ventb200.cu"]
HEADERS = ["vb_platform.h", "vb_rt.h", "vb_common.cuh", "n4_kernels.cuh", "vdp_kernels.cuh", "ci_kernels.cuh",
           "overlay_kernels.cuh", "recon_kernels.cuh", "io_kernels.cuh",
           os.path.join("..", "..", "include", "ventb200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed")
    with open(os.path.join(HERE, "lib", "ptxas.log"), "w") as f:
        f.write(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
