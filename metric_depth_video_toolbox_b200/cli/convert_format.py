"""convert_metric_depth_video_to_other_format.py front end (reference :476-962) for the dense per-frame
exports: --save_ply (decode D2 -> unproject float64 -> pose -> binary PLY per frame), --bit16 / --bit8 grey
depth video.  The sparse tracking / triangulation / Alembic / OBJ features of the script are host-side
geometry outside the per-pixel path and are refused."""
from __future__ import annotations

import argparse
import json
import os
import sys
from typing import List, Optional

import numpy as np
import torch

from .. import ops, ply, video_io
from ..geometry import compute_camera_matrix, fov_from_camera_matrix, rebase_transformations

OUT_OF_SCOPE = ("save_obj", "track_file", "strict_mask", "mask_video", "merge_close_points", "show_scene_point_clouds", "show_both_point_clouds",
                "save_alembic", "use_triangulated_points", "save_rescaled_depth", "global_align", "remove_edges")


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Convert depth video other formats like .obj or .ply or greyscale video")
    add = p.add_argument
    add("--depth_video", type=str, required=True, help="video file to use as input")
    add("--bit16", action="store_true", help="Convert depth video to a 16bit mono grayscale video file")
    add("--bit8", action="store_true", help="Convert depth video to a rgb grayscale video file")
    add("--max_depth", default=100, type=int, help="the max depth that the video uses")
    add("--save_ply", type=str, help="folder to save .ply pointcloud files in")
    add("--save_obj", type=str, help="folder to save .obj mesh files in")
    add("--color_video", type=str, help="video file to use as color input")
    add("--xfov", type=float, help="fov in deg in the x-direction, calculated from aspect ratio and yfov if not given")
    add("--yfov", type=float, help="fov in deg in the y-direction, calculated from aspect ratio and xfov if not given")
    add("--min_frames", default=-1, type=int, help="start conversion after nr of frames")
    add("--max_frames", default=-1, type=int, help="quit after max_frames nr of frames")
    add("--transformation_file", type=str, help="file with scene transformations from the aligner")
    add("--transformation_lock_frame", default=0, type=int, help="the frame that the transformation will use as a base")
    add("--remove_edges", action="store_true", help="Tries to remove edges that were not visible in the image")
    add("--track_file", type=str, help="file with 2d point tracking data")
    add("--strict_mask", default=False, action="store_true", help="Remove any points that have ever been masked out")
    add("--mask_video", type=str, help="black and white mask video for things that should not be tracked")
    add("--merge_close_points", action="store_true", help="Merges points that are very close to each other")
    add("--show_scene_point_clouds", action="store_true", help="Opens window and shows the resulting pointclouds")
    add("--show_both_point_clouds", action="store_true", help="If the viewer should show both pointclouds overlapping")
    add("--save_alembic", action="store_true", help="Save data to an alembic file")
    add("--use_triangulated_points", action="store_true", help="If the triangulated points should be used")
    add("--tringulation_min_observations", default=5, type=int, help="Nr of observations of a tracked point required")
    add("--save_rescaled_depth", action="store_true", help="Saves a video with rescaled depth")
    add("--global_align", action="store_true", help="Aligns the depth video to the triangulated depth")
    return p


def main(argv: Optional[List[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    if not os.path.isfile(args.depth_video):
        raise Exception("input video does not exist")
    for flag in OUT_OF_SCOPE:
        if getattr(args, flag):
            raise NotImplementedError(f"--{flag}: sparse / mesh-topology feature outside the dense per-frame GPU path")
    if args.color_video is not None and not os.path.isfile(args.color_video):
        raise Exception("input color_video does not exist")
    w, h, fps, total = video_io.video_info(args.depth_video)
    cam_matrix = None
    if args.save_ply is not None:
        if args.xfov is None and args.yfov is None:
            print("Either --xfov or --yfov is required.")
            return 0
        os.makedirs(args.save_ply, exist_ok=True)
    if args.xfov is not None or args.yfov is not None:
        cam_matrix = compute_camera_matrix(args.xfov, args.yfov, w, h)
        fovx, fovy = fov_from_camera_matrix(cam_matrix)
        print("Camera fovx: ", fovx, "fovy:", fovy)
    transformations = None
    if args.transformation_file is not None:
        if not os.path.isfile(args.transformation_file):
            raise Exception("input transformation_file does not exist")
        with open(args.transformation_file) as fh:
            transformations = rebase_transformations(json.load(fh), args.transformation_lock_frame)

    device = torch.device("cuda", torch.cuda.current_device())
    first = args.min_frames + 1 if args.min_frames != -1 else 0         # frames <= min_frames are skipped (:637-639)
    last = total if args.max_frames == -1 else min(total, args.max_frames + 2)  # the loop breaks only after frame max_frames + 1 (:763-766)
    grey = None
    if args.bit16 or args.bit8:
        import cv2

        out_path = args.depth_video + "_grey_depth.mkv"
        if args.bit16:
            grey = cv2.VideoWriter(filename=out_path, apiPreference=cv2.CAP_FFMPEG, fourcc=cv2.VideoWriter_fourcc(*"FFV1"), fps=fps,
                                   frameSize=(w, h), params=[cv2.VIDEOWRITER_PROP_DEPTH, cv2.CV_16U, cv2.VIDEOWRITER_PROP_IS_COLOR, 0])
        else:
            grey = cv2.VideoWriter(out_path, cv2.VideoWriter_fourcc(*"FFV1"), fps, (w, h))
    src = None if cam_matrix is None else ops.make_source(w, h, cam_matrix, args.max_depth, "D2", True, 1.0, True)
    frame_n = first
    for n, (depth_rgb, colour) in video_io.open_chunk_reader([args.depth_video, args.color_video], first, last, chunk=4):
        d = depth_rgb.to(device, non_blocking=True)
        if grey is not None:
            for frame in ops.depth_to_grey(d, args.max_depth, 16 if args.bit16 else 8, "D2").cpu().numpy():
                grey.write(frame)
        for k in range(n):
            print(f"Frame: {frame_n} {frame_n / fps}s", end="\r", file=sys.stderr)
            if args.save_ply is not None:
                pose = None if transformations is None else transformations[frame_n]
                xyz = ops.unproject(d[k], src, pose, torch.float64, cam_matrix).cpu().numpy()
                rgb = (depth_rgb if colour is None else colour)[k].cpu().numpy().reshape(-1, 3)
                ply.write_point_cloud(os.path.join(args.save_ply, f"{frame_n:07d}.ply"), xyz, rgb)
            frame_n += 1
    if grey is not None:
        grey.release()
    return 0


if __name__ == "__main__":
    sys.exit(main())
