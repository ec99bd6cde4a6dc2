"""Per-kernel roofline table from an `ncu --set full` report that holds one launch of each of the
small kernels (run in the build container):  python tools/ncu_kernels.py REPORT.ncu-rep SUBJECTS
DRAM bytes are ncu's dram__bytes_read/write; algorithmic bytes are DESIGN.md section 4's per-voxel
figures x the voxels of the launch (C2 volumes, 128x128x14)."""
import csv, io, subprocess, sys, json

PEAK = 6548.2e9
try:
    PEAK = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]) * 1e9
except Exception:
    pass

def main(rep, subjects, defects_per_subject=960):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    V = subjects * 128 * 128 * 14
    alg = {"snr_kernel": 6 * V, "vdp_stats_kernel": 5 * V, "vdp_classify4_kernel": 8 * V, "vdp_classify_kernel": 8 * V,
           "ci_pack_kernel": V * 1.125 + 4 * subjects * defects_per_subject}
    def val(r, k):
        i = hdr.index(k)
        v = float(r[i].replace(",", ""))
        u = units[i]
        return v * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    print(f"{rep}: {subjects} subjects per launch, HBM peak {PEAK/1e9:.1f} GB/s (MEASURED_PEAKS.json)")
    print("| kernel | time | DRAM read + write | DRAM GB/s (% of peak) | algorithmic bytes | algorithmic GB/s (% of peak) | issue slots busy | warp instructions |")
    print("|---|---|---|---|---|---|---|---|")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        t = val(r, "gpu__time_duration.sum")
        d = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        a = alg.get(name.split("<")[0])
        iss = float(r[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")])
        inst = float(r[hdr.index("smsp__inst_executed.sum")].replace(",", ""))
        atxt = f"{a/1e9:.2f} GB | {a/t/1e9:.0f} ({100*a/t/PEAK:.0f} %)" if a else "not HBM-bound | - "
        print(f"| `{name}` | {t*1e3:.3f} ms | {d/1e9:.2f} GB | {d/t/1e9:.0f} ({100*d/t/PEAK:.0f} %) | {atxt} | {iss:.0f} % | {inst/1e6:.0f} M |")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1776)
