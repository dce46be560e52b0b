"""The generic stereo frame loop in its steady state for ncu --cache-control none (DRAM traffic with the caches as the loop
leaves them): one warm-up call over 16 frames, then a second call whose kernels are captured."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip
w, h, n = 1920, 1080, 16
d, c = SyntheticClip(w, h, n).frames(0, 4)
d = torch.from_numpy(np.concatenate([d] * 4)).cuda(); c = torch.from_numpy(np.concatenate([c] * 4)).cuda()
rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=[5.0] * n, infill_mask=True, force_generic=True), "cuda")
sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device="cuda"); mask = torch.empty((n, h, 2 * w), dtype=torch.uint8, device="cuda")
for _ in range(2):
    rr.render_device(d, c, 0, sbs, mask)
    torch.cuda.synchronize()
print("holes", float((mask == 255).float().mean()))
