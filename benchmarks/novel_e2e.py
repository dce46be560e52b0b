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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 36
w = int(sys.argv[2]) if len(sys.argv) > 2 else 3840
h = int(sys.argv[3]) if len(sys.argv) > 3 else 2160
tmp = os.path.join(tempfile.gettempdir(), f"mdvt_novel_e2e_{n}_{w}")
shutil.rmtree(tmp, ignore_errors=True)
os.makedirs(tmp)
t0 = time.time()
depth, colour = SyntheticClip(w, h, n).frames()
for name, frames in (("depth.mkv", depth), ("colour.mkv", colour)):   # OpenCV's files (GOP 12), written on all cores
    pw = video_io.ParallelWriter(os.path.join(tmp, name), 24.0, (w, h), lanes=os.cpu_count(), block=12)
    pw.write(frames, rgb=True)
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
                  "result_writer": "gpu (mdvt_ffv1_encode_frames)" if os.environ.get("MDVT_FFV1_WRITER") == "gpu" else "host lanes (cv2.VideoWriter x cores)",
                  "synthetic_clip_write_s": round(t_gen, 2), "render_s": round(t, 2), "frames_per_s": round(n / t, 2)}))
shutil.rmtree(tmp, ignore_errors=True)
