"""Brief per-launch table of an .ncu-rep: time, DRAM bytes, L2 hit rate, issue-active, instructions, occupancy, top stalls.
    python benchmarks/ncu_brief.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
def col(v, k):
    return v[hdr.index(k)] if k in hdr else "nan"
stalls = [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for v in rows[2:]:
    f = lambda k: float(col(v, k).replace(",", "") or "nan")
    name = col(v, "Kernel Name")[:60]
    t = f("gpu__time_duration.sum"); tu = units[hdr.index("gpu__time_duration.sum")]
    rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum"); ru = units[hdr.index("dram__bytes_read.sum")]
    top = sorted(((f(k), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k in stalls), reverse=True)[:4]
    print(f"{name}\n   {t:.1f} {tu}  dram rd {rd:.1f} wr {wr:.1f} {ru}  dram% {f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f}  L2hit {f('lts__t_sector_hit_rate.pct'):.0f}%  "
          f"issue {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f}%  inst {f('smsp__inst_executed.sum')/1e6:.2f}M  warps {f('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f}%  "
          f"regs {col(v,'launch__registers_per_thread')} grid {col(v,'launch__grid_size')}\n   stalls: " + ", ".join(f"{n} {x:.1f}" for x, n in top))
