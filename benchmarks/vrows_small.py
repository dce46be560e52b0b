"""A few small frames through the virtual-source-row kernel (for compute-sanitizer memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip
for (w, h, n), convs in (((32, 42, 1), [3.0]), ((96, 40, 2), [1.0, 5.0]), ((352, 24, 2), [0.8, 2.0]), ((640, 36, 1), [0.7])):
    d, c = SyntheticClip(w, h, n, zero_fraction=0.01).frames()
    d, c = torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda()
    outs = []
    for kernel in ("vrows", "generic"):
        rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=convs, infill_mask=True, conv_kernel=kernel), "cuda")
        dep = torch.zeros((n, h, 2 * w), dtype=torch.float32, device="cuda")
        sbs, mask = rr.render_device(d, c, out_depth=dep)
        torch.cuda.synchronize()
        outs.append((sbs, mask, dep))
    print(w, h, n, "identical", bool(torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])))
