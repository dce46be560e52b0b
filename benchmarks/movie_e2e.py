"""BASELINE configs[4] in miniature at full resolution: movie_2_3D step 4 (convergence depths) + step 5 (stereo SBS +
normals-coded infill mask, the flags movie_2_3D passes) on a synthetic 1920x1080 FFV1 clip, files in, files out.
Prints one JSON line with the wall-clock stage times.  The codecs (OpenCV FFV1 decode / encode) and the TELEA inpaint
of the infill mask run on host cores exactly as in the reference; the per-pixel path runs on the GPU.

    python benchmarks/movie_e2e.py [frames=48] [--green]        (--green: green/black mask, no TELEA)
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/movie_e2e.py 96
"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from metric_depth_video_toolbox_b200 import movie_steps, sharding, video_io
from metric_depth_video_toolbox_b200.cli import stereo_rerender
from metric_depth_video_toolbox_b200.synth import SyntheticClip

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 48
green = "--green" in sys.argv
w, h = 1920, 1080
rank, world, local = sharding.init_from_env()
tmp = os.environ.get("MDVT_E2E_DIR") or os.path.join(tempfile.gettempdir(), f"mdvt_e2e_{os.environ.get('MASTER_PORT', 'single')}_{n}")
if rank == 0:
    import shutil

    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
paths = {k: os.path.join(tmp, k + ".mkv") for k in ("depth", "colour", "mask")}
t_gen = time.time()
host_inputs = "--host-inputs" in sys.argv   # inputs as the REFERENCE's tools write them (cv2.VideoWriter: 2 x 2 slices, GOP 12): host decode
if rank == 0:
    depth, colour = SyntheticClip(w, h, n).frames()
    for key, frames in (("depth", depth), ("colour", colour), ("mask", np.full((n, h, w, 3), 255, np.uint8))):
        if host_inputs:
            pw = video_io.ParallelWriter(paths[key], 24.0, (w, h), lanes=os.cpu_count(), block=12)   # OpenCV's files, GOP 12, on all cores
        else:   # inputs as THIS package's writers leave them (save_depth_video, the result writers): decoded on the device
            from metric_depth_video_toolbox_b200 import ffv1_gpu

            pw = ffv1_gpu.GpuFfv1Writer(paths[key], 24.0, (w, h), device=torch.device("cuda", local))
        for a in range(0, n, 8):
            pw.write(frames[a:a + 8], rgb=True)
        pw.close()
if world > 1:
    torch.distributed.barrier()
t_gen = time.time() - t_gen
scene = {"finished": False, "scene_video_file": paths["colour"], "depth_video_file": paths["depth"], "mask_video_file": paths["mask"],
         "xfov": 60.0, "sbs": os.path.join(tmp, "sbs.mkv")}
t0 = time.time()
movie_steps.step4_find_convergence([scene])
t4 = time.time() - t0
t0 = time.time()
if green:
    argv = movie_steps.stereo_rerender_argv(scene) + ["--green_and_black_infill_mask"]
    if os.environ.get("MDVT_E2E_CHUNK"):
        argv += ["--chunk_frames", os.environ["MDVT_E2E_CHUNK"]]
    stereo_rerender.run(stereo_rerender.build_parser().parse_args(argv), keep_process_group=True)
else:
    movie_steps.step5_render_sbs(None, [scene])
t5 = time.time() - t0
if rank == 0:
    out = paths["depth"] + "_stereo.mkv"
    assert os.path.isfile(out) and os.path.isfile(out + "_infillmask.mkv"), os.listdir(tmp)
    print(json.dumps({"workload": f"movie_2_3D steps 4+5, {w}x{h} x {n} frames, FFV1 in/out, {'green/black' if green else 'normals-coded + TELEA'} infill mask",
                      "n_gpus": world, "host_cores": os.cpu_count(),
                      "inputs": "cv2.VideoWriter files (host decode)" if host_inputs else "GpuFfv1Writer files (device decode, mdvt_ffv1_decode_frames)",
                      "result_writer": "gpu (mdvt_ffv1_encode_frames)" if video_io.gpu_ffv1_requested() else "host lanes (cv2.VideoWriter x cores)",
                      "synthetic_clip_write_s": round(t_gen, 2),
                      "step4_s": round(t4, 2), "step4_frames_per_s": round(n / t4, 1), "step5_s": round(t5, 2),
                      "step5_frames_per_s": round(n / t5, 1), "frames_per_s": round(n / (t4 + t5), 2)}))
