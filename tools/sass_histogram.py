"""SASS opcode histogram per kernel of the shipped library (run in the build container):
    python tools/sass_histogram.py [lib.so] > profiles/rNN_sass_opcodes.md
Shows which data-movement / synchronisation instructions each kernel really contains."""
import collections, os, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "vent_analysis_b200", "lib", "libventb200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for ln in txt.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", ln)
    if m and kern:
        op = m.group(1)
        hist[kern][op.split(".")[0]] += 1
        if op.startswith(("LDG", "STG", "LD.", "ST.", "ATOM", "RED", "LDS", "STS", "ATOMS", "UBLKCP", "SYNCS", "LDGSTS", "UTMA", "BAR", "REDUX", "LDL", "STL")):
            hist[kern]["~" + ".".join(op.split(".")[:2])] += 0   # keep the key for the detail column
KEY = ["LDG", "STG", "LD", "ST", "LDS", "STS", "ATOMS", "ATOM", "RED", "REDG", "UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "UTMASTG",
       "BAR", "REDUX", "SHFL", "VOTE", "LDL", "STL", "DFMA", "DADD", "DMUL", "FFMA", "MUFU"]
print(f"SASS opcode counts per kernel of `{os.path.basename(lib)}` (static instruction counts, `cuobjdump -sass`)\n")
print("| kernel | instructions | " + " | ".join(KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
for k, h in hist.items():
    tot = sum(v for o, v in h.items() if not o.startswith("~"))
    print(f"| `{k}` | {tot} | " + " | ".join(str(h.get(o, 0)) for o in KEY) + " |")
print("\nLD / ST / ATOM / RED without a suffix are GENERIC-address instructions; LDS / STS / ATOMS are shared-memory ones; "
      "UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier operations; LDL / STL = local memory (spills, local arrays).")
