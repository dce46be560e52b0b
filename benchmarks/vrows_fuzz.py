"""One-off stress of the virtual-row convergence kernel against the generic frame loop: N random geometries (the generator of
tests/test_gpu_parity.py::test_stereo_conv_vrows_random_geometries_equal_the_generic_loop with other seeds, wider frames
and taller frames so that CTAs walk many units and frames change inside a CTA); every byte, mask and depth bit must agree.

    python benchmarks/vrows_fuzz.py [cases=300] [seed=1]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from metric_depth_video_toolbox_b200 import ops
from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dev = "cuda"
taken = bad = 0
for case in range(cases):
    big = rng.random() < 0.15
    w = 32 * int(rng.integers(1, 61 if big else 25))
    h = int(rng.integers(2, 700 if big else 160))
    n = int(rng.integers(1, 6))
    xfov = float(rng.uniform(35.0, 105.0))
    yfov = None if rng.random() < 0.5 else float(rng.uniform(30.0, 90.0))
    convs = [float(np.exp(rng.uniform(np.log(0.6), np.log(60.0)))) for _ in range(n)]
    ipd = float(rng.uniform(50.0, 75.0))
    infill = bool(rng.random() < 0.7)
    depth, colour = SyntheticClip(w, h, n, seed=1000 + case, zero_fraction=0.01, n_rects=4).frames()
    common = dict(xfov=xfov, yfov=yfov, convergence_depths=convs, pupillary_distance=ipd, master_xfov=float(rng.uniform(40.0, 60.0)), infill_mask=infill)
    if rng.random() < 0.3:
        common.update(xfov=None, yfov=None, xfovs=[float(xfov + 3.0 * k) for k in range(n)])
    probe = StereoRerenderer(StereoParams(w, h, **common), dev)
    host = ops.conv_frames_packed(*probe.packed_cameras(0, n), probe.p.near)
    if not ops.conv_vrows_supported(host, w, h):
        continue
    taken += 1
    d, c = torch.from_numpy(depth).to(dev), torch.from_numpy(colour).to(dev)
    outs = []
    for kernel in ("vrows", "generic"):
        rr = StereoRerenderer(StereoParams(w, h, conv_kernel=kernel, **common), dev)
        out_depth = torch.full((n, h, 2 * w), -1.0, dtype=torch.float32, device=dev)
        sbs, mask = rr.render_device(d, c, out_depth=out_depth)
        outs.append((sbs, mask, out_depth))
    (sa, ma, da), (sb, mb, db) = outs
    ok = torch.equal(sa, sb) and ((ma is None and mb is None) or torch.equal(ma, mb)) and torch.equal(da.view(torch.int32), db.view(torch.int32))
    if not ok:
        bad += 1
        print("MISMATCH", case, w, h, n, xfov, yfov, convs, ipd, infill)
torch.cuda.synchronize()
print(f"{cases} cases, {taken} inside the kernel's limits, {bad} mismatches")
