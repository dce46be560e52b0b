"""Convergence stereo through the virtual-row kernel at 1280x720 (39 KB of shared memory per CTA: room for a fifth CTA per SM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip
w, h, n = 1280, 720, 64
d, c = SyntheticClip(w, h, n).frames(0, 4)
d = torch.from_numpy(np.concatenate([d] * 16)).cuda(); c = torch.from_numpy(np.concatenate([c] * 16)).cuda()
sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device="cuda"); mask = torch.empty((n, h, 2 * w), dtype=torch.uint8, device="cuda")
rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=[5.0 + 0.02 * f for f in range(n)], infill_mask=True, conv_kernel="vrows"), "cuda")
for _ in range(3): rr.render_device(d, c, 0, sbs, mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): rr.render_device(d, c, 0, sbs, mask)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 / n * 1e3
print(f"720p convergence stereo: {us:.2f} us/frame = {us / (w * h) * 1e6:.2f} ps/px (1080p-equivalent {us * 2.25:.2f} us)")
