"""One mdvt_ffv1_encode_frames call over 8 frames of 3840x1080 (for ncu: -k regex:ffv1_encode -c 1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffv1_gpu_bench import frames_like  # noqa: E402
from metric_depth_video_toolbox_b200 import ffv1_gpu  # noqa: E402

dev = torch.device("cuda:0")
d = torch.from_numpy(frames_like(3840, 1080, 8)).to(dev)
enc = ffv1_gpu.Ffv1Encoder(3840, 1080, dev, max_frames=8)
enc.encode_device(d)
torch.cuda.synchronize()
print("bytes", int(enc.offsets[8 * enc.per_frame]))
