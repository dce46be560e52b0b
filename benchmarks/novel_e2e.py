"""BASELINE configs[2] through files: `3d_view_depthfile.py --render` on a synthetic 3840x2160 FFV1 clip (depth + colour in,
`_render.mkv` out), wall clock.  Decode / encode are OpenCV FFV1 on host threads, the per-pixel path is on the GPU.

    python benchmarks/novel_e2e.py [frames=36] [width=3840] [height=2160]
"""
import importlib
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from metric_depth_video_toolbox_b200 import video_io
from metric_depth_video_toolbox_b200.cli import view_depthfile
from metric_depth_video_toolbox_b200.synth import SyntheticClip

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if len(args) > 0 else 36
w = int(args[1]) if len(args) > 1 else 3840
h = int(args[2]) if len(args) > 2 else 2160
tmp = os.path.join(tempfile.gettempdir(), f"mdvt_novel_e2e_{n}_{w}")
shutil.rmtree(tmp, ignore_errors=True)
os.makedirs(tmp)
t0 = time.time()
depth, colour = SyntheticClip(w, h, n).frames()
host_inputs = "--host-inputs" in sys.argv
for name, frames in (("depth.mkv", depth), ("colour.mkv", colour)):
    if host_inputs:   # OpenCV's files (GOP 12, what the reference's tools write), on all cores: decoded on host threads
        pw = video_io.ParallelWriter(os.path.join(tmp, name), 24.0, (w, h), lanes=os.cpu_count(), block=12)
    else:             # this package's writer (what save_depth_video / the result writers leave): decoded on the device
        from metric_depth_video_toolbox_b200 import ffv1_gpu

        pw = ffv1_gpu.GpuFfv1Writer(os.path.join(tmp, name), 24.0, (w, h))
    for a in range(0, n, 4):
        pw.write(frames[a:a + 4], rgb=True)
    pw.close()
t_gen = time.time() - t0
del depth, colour
t0 = time.time()
view_depthfile.main(["--depth_video", os.path.join(tmp, "depth.mkv"), "--color_video", os.path.join(tmp, "colour.mkv"), "--xfov", "60",
                     "--render", "--x", "2", "--y", "2", "--z", "-4"])
t = time.time() - t0
out = os.path.join(tmp, "depth.mkv_render.mkv")
assert video_io.video_info(out)[3] == n
print(json.dumps({"workload": f"3d_view_depthfile.py --render, {w}x{h} x {n} frames, FFV1 in/out", "host_cores": os.cpu_count(),
                  "inputs": "cv2.VideoWriter files (host decode)" if host_inputs else "GpuFfv1Writer files (device decode)",
                  "result_writer": "gpu (mdvt_ffv1_encode_frames)" if video_io.gpu_ffv1_requested() else "host lanes (cv2.VideoWriter x cores)",
                  "synthetic_clip_write_s": round(t_gen, 2), "render_s": round(t, 2), "frames_per_s": round(n / t, 2)}))
shutil.rmtree(tmp, ignore_errors=True)
