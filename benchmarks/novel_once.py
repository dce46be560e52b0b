"""One small chunk of the 4K novel-view path (config 3) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metric_depth_video_toolbox_b200.novel_view import NovelViewParams, NovelViewRenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip
w, h, n = 3840, 2160, 3
d, c = SyntheticClip(w, h, n).frames(0, n)
d, c = torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda()
nv = NovelViewRenderer(NovelViewParams(w, h, 60, None, 100), "cuda")
rgb, mask = nv.render_device(d, c)
torch.cuda.synchronize()
print("holes", float((mask == 255).float().mean()))
