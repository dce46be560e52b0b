#!/bin/bash
# round 2: ncu --set full of the final virtual-row kernel (32 frames of 1080p in one launch)
mkdir -p gpurun_out
cat > /tmp/vrows32.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip
w, h, n = 1920, 1080, 32
d, c = SyntheticClip(w, h, n).frames(0, 4)
d = torch.from_numpy(np.concatenate([d] * 8)).cuda(); c = torch.from_numpy(np.concatenate([c] * 8)).cuda()
rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=[5.0 + 0.02 * f for f in range(n)], infill_mask=True, conv_kernel="vrows"), "cuda")
for _ in range(3):
    sbs, mask = rr.render_device(d, c)
torch.cuda.synchronize()
print("holes", float((mask == 255).float().mean()))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vrows -s 2 -c 1 -f -o gpurun_out/r02_vrows_v13_32f python /tmp/vrows32.py > gpurun_out/r02_vrows_ncu_v13.log 2>&1; tail -2 gpurun_out/r02_vrows_ncu_v13.log
