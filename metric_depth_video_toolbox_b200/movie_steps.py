"""The two steps of movie_2_3D.py that run the dense per-frame path, with the reference's signatures
(movie_2_3D.py:408-452), executed in-process on the GPU instead of as `python <script>` subprocesses:

    step4_find_convergence(scene_video_files)        per-scene convergence depth lists (find_convergence_depth.py)
    step5_render_sbs(args, scene_video_files)        per-scene stereo SBS + infill-mask videos (stereo_rerender.py)

The reference parallelises step 5 over scenes with up to `args.parallel` subprocesses.  Here every scene is rendered
by all GPUs of the job at once (frames sharded across the torchrun ranks inside the stereo front end), scenes one
after the other, so `args.parallel` is accepted and ignored.  The other steps of movie_2_3D (scene detection, depth
models, masks, learned infill, ffmpeg muxing) are outside the dense per-frame path."""
from __future__ import annotations

import os
from typing import Dict, List

from . import sharding
from .cli import find_convergence_depth, stereo_rerender


from .movie_plan import is_valid_video  # noqa: E402,F401  (movie_2_3D.py:62-67: exists and holds at least 2 KB)


def _barrier():
    if sharding.world()[1] > 1:
        import torch.distributed as dist

        dist.barrier()


def step4_find_convergence(scene_video_files: List[Dict]) -> None:
    """Compute convergence depths for each scene (movie_2_3D.py:408-419).  One small reduction kernel per frame; under
    torchrun rank 0 does it while the others wait."""
    print("Step four: find convergence depth for focus point")
    rank, _ = sharding.world()
    for scene in scene_video_files:
        if scene["finished"]:
            continue
        assert is_valid_video(scene["mask_video_file"]), "Could find valid mask video: " + scene["mask_video_file"]
        scene["convergence_file"] = scene["depth_video_file"] + "_convergence_depths.json"
        if rank == 0 and not os.path.exists(scene["convergence_file"]):
            find_convergence_depth.main(["--depth_video", scene["depth_video_file"], "--mask_video", scene["mask_video_file"]])
    _barrier()


def stereo_rerender_argv(scene: Dict) -> List[str]:
    """The command line movie_2_3D builds for one scene (movie_2_3D.py:431-445), as an argv list."""
    argv = ["--color_video", scene["scene_video_file"]]
    if scene.get("convergence", True):
        argv += ["--convergence_file", scene["convergence_file"]]
    if scene.get("xfov") is not None:
        argv += ["--xfov", str(scene["xfov"])]
    else:
        argv += ["--xfov_file", scene["xfovs_file"]]
    argv += ["--depth_video", scene["depth_video_file"]]
    if scene.get("infill", True):
        argv.append("--infill_mask")
    return argv


def step5_render_sbs(args, scene_video_files: List[Dict]) -> None:
    """Render stereo (SBS) frames (movie_2_3D.py:422-452)."""
    print("Step five: render SBS frames")
    for scene in scene_video_files:
        if os.path.exists(scene["sbs"]) or scene["finished"]:
            continue
        stereo_rerender.run(stereo_rerender.build_parser().parse_args(stereo_rerender_argv(scene)), keep_process_group=True)
    _barrier()
