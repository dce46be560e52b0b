"""A few frames of convergence stereo (the fused target-row kernel, what movie_2_3D selects) at 1080p for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip
w, h, n = 1920, 1080, 8
kernel = sys.argv[1] if len(sys.argv) > 1 else "rows"   # rows | vrows | generic
conv = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
d, c = SyntheticClip(w, h, n).frames(0, n)
d, c = torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda()
rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=[conv] * n, infill_mask=True, conv_kernel=kernel), "cuda")
for _ in range(2):
    sbs, mask = rr.render_device(d, c)
torch.cuda.synchronize()
print("holes", float((mask == 255).float().mean()))
