"""Per-kernel timing of the infill-mask renderer (movie_2_3D's step-5 mode) at 1080p: the stereo render, the mesh edge test,
the edge-point splat and the mask painting, one frame at a time as InfillMaskRenderer.render_device issues them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metric_depth_video_toolbox_b200 import geometry as geo, ops
from metric_depth_video_toolbox_b200.infill import InfillMaskRenderer
from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip

w, h, n = 1920, 1080, 12
d, c = SyntheticClip(w, h, n).frames(0, n)
d, c = torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda()
p = StereoParams(w, h, xfov=60.0, convergence_depths=[5.0] * n, infill_mask=True)
rr = StereoRerenderer(p, "cuda")
inf = InfillMaskRenderer(rr)
sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device="cuda")
mask = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device="cuda")


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps


for code in (False, True):
    ms, wall = timed(lambda: inf.render_device(d, c, 0, sbs, mask, code, True, None))
    print(f"InfillMaskRenderer.render_device code_normals={code}: {ms / n:.3f} ms/frame GPU, {wall / n:.3f} ms/frame wall")
K = geo.compute_camera_matrix(60.0, None, w, h)
src = ops.make_source(w, h, K, 100, "D1", True, geo.master_fov_depth_scale(45.0, 60.0), True)
flags = torch.empty((h, w), dtype=torch.uint8, device="cuda"); normals = torch.empty((h, w, 3), dtype=torch.float64, device="cuda")
zb = ops.new_zbuf(1, w, h, "cuda")[0]
holes = torch.zeros((h, 2 * w), dtype=torch.uint8, device="cuda")
view = rr.views_of(0)[0]
for name, fn in (("stereo render (12 frames)", lambda: rr.render_device(d, c, 0, sbs, None, None, mask_rgb=False)),
                 ("edge_vertices", lambda: ops.edge_vertices(d[0], src, K, True, flags=flags, normals=normals)),
                 ("edge_splat", lambda: ops.edge_splat(d[0], src, K, flags, view.M, K, w, h, zb)),
                 ("edge_resolve", lambda: ops.edge_resolve(zb, d[0], src, K, normals, view.M, c[0], holes[:, :w], mask[0, :, :w], sbs[0, :, :w], (0, 255, 0), True)),
                 ("views_of + make_source (host only)", lambda: (rr.views_of(0), ops.make_source(w, h, K, 100, "D1", True, 1.0, True)))):
    ms, wall = timed(fn, 5)
    print(f"{name}: {ms:.3f} ms GPU, {wall:.3f} ms wall")
