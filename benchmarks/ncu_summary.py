"""Summarise an .ncu-rep (raw page) into the few numbers the roofline needs; writes JSON to stdout.
    python benchmarks/ncu_summary.py gpurun_out/x.ncu-rep [frames_per_launch]"""
import csv, json, subprocess, sys

rep = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second"]
out = []
for vals in rows[2:]:
    d = {}
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            d[k] = vals[i] + (" " + units[i] if units[i] else "")
    out.append(d)
print(json.dumps({"report": rep, "frames_per_launch": frames, "launches": out}, indent=1))
