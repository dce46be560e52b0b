#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_dropin.py -q -x -k "novel" 2>&1 | tail -3
for chunk in 1 2 4; do echo -n "novel e2e chunk=$chunk: "; MDVT_HOST_CHUNK=$chunk timeout 300 python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from metric_depth_video_toolbox_b200.novel_view import NovelViewParams, NovelViewRenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip
w, h, n = 3840, 2160, 16
d, c = SyntheticClip(w, h, n).frames(0, 4)
hd = torch.from_numpy(np.concatenate([d] * 4)).pin_memory(); hc = torch.from_numpy(np.concatenate([c] * 4)).pin_memory()
nv = NovelViewRenderer(NovelViewParams(w, h, 60.0, None, 100), torch.device("cuda:0"))
out = torch.empty((n, h, w, 3), dtype=torch.uint8, pin_memory=True)
nv.render_host(hd, hc, out); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): nv.render_host(hd, hc, out)
torch.cuda.synchronize()
print(round(3 * n / (time.perf_counter() - t0)), "frames/s")
PY
done
