"""Per-kernel CUDA-event timing of the 4K novel-view path (config 3): centroid reduction, K1+K2 splat, K3 resolve,
each timed alone over `reps` back-to-back launches on distinct frames (inputs > L2 in total)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from metric_depth_video_toolbox_b200 import ops
from metric_depth_video_toolbox_b200.novel_view import NovelViewParams, NovelViewRenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


w, h, n = 3840, 2160, 8
if len(sys.argv) > 1 and sys.argv[1] == "1080":
    w, h, n = 1920, 1080, 16
d, c = SyntheticClip(w, h, n).frames(0, 2)
d = torch.from_numpy(np.concatenate([d] * (n // 2))).cuda()
c = torch.from_numpy(np.concatenate([c] * (n // 2))).cuda()
nv = NovelViewRenderer(NovelViewParams(w, h, 60, None, 100), "cuda")
centres = nv.centroids(d)
src_c = ops.make_source(w, h, nv.K, 100, "D1", True, 1.0, True)
src = ops.make_source(w, h, nv.K, 100, "D1", True, 1.0, False)
sums = torch.empty((n, 4 + ops._lib.REDUCE_SCRATCH_DOUBLES), dtype=torch.float64, device="cuda")
zb = ops.new_zbuf(1, w, h, "cuda")
rgb = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
mask = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
views = [[nv.view_of(k, centres[k])] for k in range(n)]
flags = ops.FLAG_RESET_ZBUF if hasattr(ops, "FLAG_RESET_ZBUF") else 0


def centroid():
    for k in range(n):
        ops.centroid_sums(d[k], src_c, nv.K, None, out=sums[k])


def splat():
    for k in range(n):
        ops.project_splat(d[k], src, views[k], w, h, zb)


def resolve():
    for k in range(n):
        ops.resolve(zb[0], c[k], (255, 255, 255), (255, 255, 255), 0, rgb[k], mask[k])


def whole():
    nv.render_device(d, c, 0, rgb, mask)


def whole_hostcam():
    nv.render_device_hostcam(d, c, 0, rgb, mask)


for name, fn in (("centroid", centroid), ("splat (accumulating zbuf)", splat), ("resolve (no reset)", resolve), ("whole render_device (one call, device camera)", whole), ("whole render_device_hostcam (v1)", whole_hostcam)):
    ms = timed(fn)
    print(f"{w}x{h} {name}: {ms / n * 1e3:.1f} us/frame")
print("holes", float((mask == 255).float().mean()))
