"""profiles/n4_counters.json from an `ncu --set full` capture of ONE n4_subject_kernel launch (run in
the build container):   python tools/n4_counters.py REPORT.ncu-rep RUNINFO.json OUT.json
RUNINFO.json is written by `VB_DUMP_INFO=... python tools/run_stage.py n4 B 1` (the profiled command)."""
import csv, io, json, subprocess, sys, os

rep, info, out = sys.argv[1:4]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
row = [r for r in rows[2:] if "n4_subject_kernel" in r[hdr.index("Kernel Name")]][0]
SC = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
def val(k):
    i = hdr.index(k)
    return float(row[i].replace(",", "")) * SC.get(units[i], 1.0)
j = json.load(open(info))
j.update({
    "source": os.path.basename(rep) + " (ncu --set full --clock-control none, one launch)",
    "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
    "warp_instructions": val("smsp__inst_executed.sum"),
    "smem_wavefronts": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "smem_bank_conflicts": val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    "duration_ms_under_ncu": val("gpu__time_duration.sum") * 1e3,
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "ipc": val("sm__inst_executed.avg.per_cycle_active"),
    "registers_per_thread": val("launch__registers_per_thread"),
    "block_size": val("launch__block_size"),
})
json.dump(j, open(out, "w"), indent=1)
print(json.dumps(j))
