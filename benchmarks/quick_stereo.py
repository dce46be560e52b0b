"""Quick kernel-only timing of the fused stereo kernel (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metric_depth_video_toolbox_b200 import ops
from metric_depth_video_toolbox_b200.synth import SyntheticClip

w, h = 1920, 1080
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
extra = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0
clip = SyntheticClip(w, h, n)
t0 = time.time()
base_d, base_c = clip.frames(0, min(n, 8))
reps = (n + len(base_d) - 1) // len(base_d)
d = torch.from_numpy(np.concatenate([base_d] * reps)[:n]).cuda()
c = torch.from_numpy(np.concatenate([base_c] * reps)[:n]).cuda()
print(f"synth {time.time()-t0:.1f}s", flush=True)
consts = torch.from_numpy(ops.stereo_frame_constants(60.0, w, 100, 63, 45.0)[None]).cuda()
sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device="cuda")
mask = torch.empty((n, h, 2 * w), dtype=torch.uint8, device="cuda")
for _ in range(3):
    ops.stereo_rows(d, c, consts, (0, 255, 0), (0, 0, 0), ops.FLAG_BG_COLLIDE | extra, sbs, mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 5
for _ in range(K):
    ops.stereo_rows(d, c, consts, (0, 255, 0), (0, 0, 0), ops.FLAG_BG_COLLIDE | extra, sbs, mask)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
bytes_ = n * w * h * 14
print(f"stereo_rows {n} frames: {ms:.3f} ms/launch  {ms/n*1e3:.2f} us/frame  {n/ms*1e3:.0f} frames/s  {bytes_/ms/1e6:.0f} GB/s  ({bytes_/ms/1e6/6454:.3f} of measured HBM peak)")
print("hole fraction", (mask == 255).float().mean().item())
