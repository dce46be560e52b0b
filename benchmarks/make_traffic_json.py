"""profiles/stereo_rows_traffic.json from an ncu capture of the headline kernel (one launch over `frames` frames): DRAM bytes
per frame, the kernel's name as ncu reports it, and the SHA-256 of the kernel's source text, so that bench.py can say whether
the capture belongs to the code that is running.

    ncu --set full --clock-control none -k regex:stereo_rows -c 1 -o gpurun_out/rows python benchmarks/quick_stereo.py 300
    python benchmarks/make_traffic_json.py gpurun_out/rows.ncu-rep 300 profiles/rNN_stereo_rows_300f.ncu-rep-summary
"""
import csv, hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ["metric_depth_video_toolbox_b200/csrc/mdvt_stereo_rows.cu", "metric_depth_video_toolbox_b200/csrc/mdvt_common.cuh"]


VROWS_SOURCES = ["metric_depth_video_toolbox_b200/csrc/mdvt_stereo_vrows.cu", "metric_depth_video_toolbox_b200/csrc/mdvt_common.cuh"]


def source_sha256(sources=None) -> str:
    h = hashlib.sha256()
    for rel in (sources or KERNEL_SOURCES):
        with open(os.path.join(ROOT, rel), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


if __name__ == "__main__":
    rep, frames = sys.argv[1], int(sys.argv[2])
    label = sys.argv[3] if len(sys.argv) > 3 else rep
    if "--vrows" in sys.argv:   # the capture is of the virtual-source-row kernel (profiles/stereo_vrows_traffic.json)
        KERNEL_SOURCES[:] = VROWS_SOURCES
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def num(key):
        i = hdr.index(key)
        v = float(vals[i].replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units[i], 1.0)

    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    out = {"dram_bytes_per_frame": (rd + wr) / frames, "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
           "frames_per_launch": frames, "algorithmic_bytes_per_frame": 14 * 1920 * 1080, "kernel": vals[hdr.index("Kernel Name")],
           "kernel_source_sha256": source_sha256(), "kernel_sources": KERNEL_SOURCES,
           "source": f"{label} (ncu --set full --clock-control none, one {frames}-frame launch, dram__bytes_read.sum + dram__bytes_write.sum)"}
    print(json.dumps(out, indent=1))
