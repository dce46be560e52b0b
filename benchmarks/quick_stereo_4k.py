"""Row kernel at 3840x2160 (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metric_depth_video_toolbox_b200 import ops
from metric_depth_video_toolbox_b200.synth import SyntheticClip
w, h, n = 3840, 2160, 60
bd, bc = SyntheticClip(w, h, n).frames(0, 3)
d = torch.from_numpy(np.concatenate([bd] * 20)[:n]).cuda(); c = torch.from_numpy(np.concatenate([bc] * 20)[:n]).cuda()
consts = torch.from_numpy(ops.stereo_frame_constants(60.0, w, 100, 63, 45.0)[None]).cuda()
sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device="cuda"); mask = torch.empty((n, h, 2 * w), dtype=torch.uint8, device="cuda")
for _ in range(3): ops.stereo_rows(d, c, consts, (0, 255, 0), (0, 0, 0), ops.FLAG_BG_COLLIDE, sbs, mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.stereo_rows(d, c, consts, (0, 255, 0), (0, 0, 0), ops.FLAG_BG_COLLIDE, sbs, mask)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"stereo_rows 4K {n} frames: {ms/n*1e3:.2f} us/frame  {n/ms*1e3:.0f} frames/s  {n*w*h*14/ms/1e6:.0f} GB/s ({n*w*h*14/ms/1e6/6454:.3f} of HBM peak)")
