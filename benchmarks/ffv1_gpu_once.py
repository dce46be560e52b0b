"""One mdvt_ffv1_encode_frames + one mdvt_ffv1_decode_frames call over a batch of 3840x1080 frames (for ncu:
-k regex:"ffv1_encode|ffv1_decode" -c 2).  MDVT_FFV1_ONCE_MODEL = 0 | 1 (context model), MDVT_FFV1_ONCE_FRAMES = batch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffv1_gpu_bench import frames_like  # noqa: E402
from metric_depth_video_toolbox_b200 import ffv1_gpu  # noqa: E402

model = int(os.environ.get("MDVT_FFV1_ONCE_MODEL", "1"))
n = int(os.environ.get("MDVT_FFV1_ONCE_FRAMES", "8"))
dev = torch.device("cuda:0")
d = torch.from_numpy(frames_like(3840, 1080, n)).to(dev)
enc = ffv1_gpu.Ffv1Encoder(3840, 1080, dev, max_frames=n, context_model=model)
packets = enc.encode(d)
dec = ffv1_gpu.Ffv1Decoder.for_config(enc.config, 3840, 1080, dev, max_frames=n)
out = dec.decode(packets)
torch.cuda.synchronize()
print("bytes", sum(len(p) for p in packets), "identical", bool(torch.equal(out, d)))
