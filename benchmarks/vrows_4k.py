"""Convergence stereo at 3840x2160: the virtual-row kernel against the generic two-lane loop (same bytes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip
w, h, n = 3840, 2160, 12
d, c = SyntheticClip(w, h, n).frames(0, 3)
d = torch.from_numpy(np.concatenate([d] * 4)).cuda(); c = torch.from_numpy(np.concatenate([c] * 4)).cuda()
sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device="cuda"); mask = torch.empty((n, h, 2 * w), dtype=torch.uint8, device="cuda")
ref = None
for kernel in ("vrows", "generic"):
    rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=[5.0 + 0.02 * f for f in range(n)], infill_mask=True, conv_kernel=kernel), "cuda")
    for _ in range(2): rr.render_device(d, c, 0, sbs, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): rr.render_device(d, c, 0, sbs, mask)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 / n * 1e3
    same = "" if ref is None else f"  identical to vrows: {bool(torch.equal(ref[0], sbs) and torch.equal(ref[1], mask))}"
    if ref is None: ref = (sbs.clone(), mask.clone())
    print(f"4K convergence stereo, {kernel}: {us:.1f} us/frame = {14 * w * h / us / 1e3:.0f} GB/s algorithmic ({14 * w * h / us / 1e3 / 6454:.3f} of HBM peak){same}")
