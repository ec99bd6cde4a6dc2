"""TEST INFRASTRUCTURE.  Compiles the product's kernel sources (vent_analysis_b200/csrc)
for the HOST with -DVB_EMULATE: one OS thread per CUDA thread, barriers for __syncthreads and
warp collectives (see csrc/vb_platform.h).  The resulting library exports the same C ABI and
lets ``pytest -m "not gpu"`` check the kernel logic against the oracle in the GPU-less build
container.  It is never loaded by the product package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "vent_analysis_b200", "csrc")
SO = os.path.join(HERE, "_build", "libventb200_emu.so")
# one OS thread per CUDA thread: the emulation uses smaller blocks than the product build (the
# kernels only assume warp-multiple block sizes), which keeps the CPU suite to a few minutes
SMALL_BLOCKS = ["-DVB_N4_THREADS=256", "-DVB_N4_CHAIN_WARPS=4", "-DVB_VDP_THREADS=256", "-DVB_CLS_THREADS=128"]


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "ventb200.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(s) <= os.path.getmtime(SO) for s in srcs):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-DVB_EMULATE", *SMALL_BLOCKS, "-ffp-contract=off",
           "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable", "-x", "c++", os.path.join(CSRC, "ventb200.cu"),
           "-o", SO]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force=True))
