"""The generic stereo loop at several frame sizes (development probe: how the resolve's cost per pixel depends on the size of
the z-buffer planes).  MDVT_DEBUG=1 switches the splat off, =2 the resolve."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip
for (w, h) in ((640, 360), (960, 540), (1280, 720), (1920, 1080), (2560, 1440)):
    n = max(8, int(32 * 1920 * 1080 / (w * h)))
    d, c = SyntheticClip(w, h, 4).frames(0, 4)
    reps = (n + 3) // 4
    d = torch.from_numpy(np.concatenate([d] * reps)[:n]).cuda(); c = torch.from_numpy(np.concatenate([c] * reps)[:n]).cuda()
    sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device="cuda"); mask = torch.empty((n, h, 2 * w), dtype=torch.uint8, device="cuda")
    rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=[5.0] * n, infill_mask=True, force_generic=True), "cuda")
    for _ in range(2): rr.render_device(d, c, 0, sbs, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): rr.render_device(d, c, 0, sbs, mask)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 4 / n * 1e3
    print(f"{w}x{h} ({n} frames, planes {2 * w * h * 8 / 1e6:.1f} MB per set): {us:.2f} us/frame = {us / (w * h) * 1e6:.2f} ps/px")
