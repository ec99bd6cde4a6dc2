"""Tuning builds of the same sources with other compile-time constants:
    python tools/build_variant.py NAME -DVB_N4_THREADS=768 ...   ->  vent_analysis_b200/lib/variants/NAME.so
Run with VENTB200_LIB=<that path> (same C ABI)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vent_analysis_b200 import build
name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(build.HERE, "lib", "variants", name + ".so")
os.makedirs(os.path.dirname(out), exist_ok=True)
cmd = [build.nvcc()] + build.NVCC_FLAGS + flags + ["-o", out, os.path.join(build.CSRC, "ventb200.cu")]
r = subprocess.run(cmd, capture_output=True, text=True)
log = r.stdout + r.stderr
if r.returncode:
    print(log); raise SystemExit(1)
import re
m = re.search(r"n4_subject_kernel.*?\n.*?\n\s+(.*spill loads)\n.*?Used (\d+) registers", log, re.S)
print(name, flags, "->", m.group(1) if m else "?", "| regs", m.group(2) if m else "?")
