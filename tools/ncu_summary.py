"""Summarise ncu reports (run in the build container): launch list shares, key raw metrics and
per-function instruction / stall-sample shares of a --set full capture."""
import csv, io, re, subprocess, sys, os, collections

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        try: v = float(r[iv].replace(",", ""))
        except ValueError: continue
        k = r[ik].split("(")[0]
        c = tot.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += v
    unit = rows[1][hdr.index("Metric Unit")]
    s = sum(v for _, v in tot.values())
    out = [f"launch list {os.path.basename(path)}: {sum(c for c,_ in tot.values())} launches, total {s:.0f} {unit}"]
    for k, (c, v) in sorted(tot.items(), key=lambda x: -x[1][1]):
        out.append(f"  {k:28s} launches {c:4d}  time {v:14.0f} {unit}  share {100*v/s:5.1f}%")
    return "\n".join(out)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = [f"raw metrics {os.path.basename(rep)} (kernel {vals[hdr.index('Kernel Name')]})"]
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS: out.append(f"  {h} = {v} {u}")
    return "\n".join(out)

def by_function(rep, src):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    lines = open(src).read().split("\n")
    starts = []
    for i, l in enumerate(lines, 1):
        if l.startswith("VB_DEV") or l.startswith("VB_KERNEL"):
            m = re.search(r"(n4_\w+|vb_\w+|ci_\w+|vdp_\w+|snr_\w+|border_\w+)\s*\(", l)
            if m: starts.append((i, m.group(1)))
    def fn(line):
        name = "?"
        for s, n in starts:
            if s <= line: name = n
        return name
    cur, inst, samp = None, {}, {}
    for r in csv.reader(io.StringIO(txt)):
        if len(r) >= 2 and r[0] == "File Path": cur = os.path.basename(r[1]); continue
        if len(r) < 10 or not r[0].isdigit(): continue
        try: ni, ns = int(r[7]), int(r[6])
        except ValueError: continue
        k = fn(int(r[0])) if cur == os.path.basename(src) else cur
        inst[k] = inst.get(k, 0) + ni; samp[k] = samp.get(k, 0) + ns
    ti, ts = sum(inst.values()), sum(samp.values())
    out = [f"per-function shares {os.path.basename(rep)}: {ti} warp instructions, {ts} stall samples"]
    for k, v in sorted(inst.items(), key=lambda x: -x[1])[:18]:
        out.append(f"  {k:28s} instructions {100*v/ti:5.1f}%  samples {100*samp[k]/ts:5.1f}%")
    return "\n".join(out)

if __name__ == "__main__":
    for a in sys.argv[1:]:
        if a.endswith(".csv"): print(launches(a))
        elif a.endswith(".ncu-rep"):
            print(raw(a))
            print(by_function(a, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vent_analysis_b200", "csrc", "n4_kernels.cuh")))
        print()
