"""Kernel-only timing of the generic path (development aid): (a) stereo with a convergence rotation at 1080p
(K1+K2 for two views + two K3 resolves per frame), (b) the 4K novel view of config 3 (centroid + K1+K2 + K3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from metric_depth_video_toolbox_b200.novel_view import NovelViewParams, NovelViewRenderer
from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def clip(w, h, n, distinct=4):
    d, c = SyntheticClip(w, h, n).frames(0, distinct)
    reps = (n + distinct - 1) // distinct
    return (torch.from_numpy(np.concatenate([d] * reps)[:n]).cuda(), torch.from_numpy(np.concatenate([c] * reps)[:n]).cuda())


which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("stereo", "both", "conv"):
    w, h, n = 1920, 1080, 32
    d, c = clip(w, h, n)
    rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=[5.0] * n, infill_mask=True, conv_kernel=True), "cuda")
    sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device="cuda")
    mask = torch.empty((n, h, 2 * w), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: rr.render_device(d, c, 0, sbs, mask), 3)
    px = w * h * n
    print(f"convergence stereo, fused target-row kernel, 1080p: {ms / n * 1e3:.1f} us/frame  {n / ms * 1e3:.0f} frames/s  {px * 14 / ms / 1e6:.0f} GB/s algorithmic "
          f"({px * 14 / ms / 1e6 / 6454:.3f} of HBM peak)  holes {float((mask == 255).float().mean()):.4f}")
if which in ("stereo", "both", "conv", "vrows"):
    w, h, n = 1920, 1080, 32
    d, c = clip(w, h, n)
    sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device="cuda")
    mask = torch.empty((n, h, 2 * w), dtype=torch.uint8, device="cuda")
    for conv in (5.0, 2.0, 1.0):
        rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=[conv + 0.02 * f for f in range(n)], infill_mask=True, conv_kernel="vrows"), "cuda")
        ms = timed(lambda: rr.render_device(d, c, 0, sbs, mask), 5)
        px = w * h * n
        print(f"convergence stereo at {conv} m, virtual-source-row kernel, 1080p: {ms / n * 1e3:.2f} us/frame  {n / ms * 1e3:.0f} frames/s  "
              f"{px * 14 / ms / 1e6:.0f} GB/s algorithmic ({px * 14 / ms / 1e6 / 6454:.3f} of HBM peak)  holes {float((mask == 255).float().mean()):.4f}")
if which in ("stereo", "both", "posed"):
    w, h, n = 1920, 1080, 32
    d, c = clip(w, h, n)
    rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=[5.0] * n, infill_mask=True, force_generic=True), "cuda")
    sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device="cuda")
    mask = torch.empty((n, h, 2 * w), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: rr.render_device(d, c, 0, sbs, mask), 3)
    px = w * h * n
    print(f"generic stereo (K1+K2 two views, 2 x K3; what a pose file selects) 1080p: {ms / n * 1e3:.1f} us/frame  {n / ms * 1e3:.0f} frames/s  "
          f"{px * 14 / ms / 1e6:.0f} GB/s algorithmic ({px * 14 / ms / 1e6 / 6454:.3f} of HBM peak)")
if which in ("novel", "both"):
    w, h, n = 3840, 2160, 24
    d, c = clip(w, h, n, 4)
    nv = NovelViewRenderer(NovelViewParams(w, h, 60, None, 100), "cuda")
    rgb = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
    mask = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: nv.render_device(d, c, 0, rgb, mask), 3)
    px = w * h * n
    print(f"novel view 4K: {ms / n * 1e3:.1f} us/frame  {n / ms * 1e3:.0f} frames/s  {px * 10 / ms / 1e6:.0f} GB/s algorithmic "
          f"({px * 10 / ms / 1e6 / 6454:.3f} of HBM peak)  holes {float((mask == 255).float().mean()):.4f}")
