"""3d_view_depthfile.py front end (reference :18-263): `--render` writes `<depth_video>_render.mkv` from a
look-at camera aimed at the per-frame vertex centroid.  The interactive Open3D window (no --render) and
--draw_frame are GUI features outside the per-pixel path and are refused with a message."""
from __future__ import annotations

import argparse
import json
import os
import sys
from typing import List, Optional

import numpy as np
import torch

from .. import video_io
from ..geometry import compute_camera_matrix, fov_from_camera_matrix, rebase_transformations
from ..novel_view import NovelViewParams, NovelViewRenderer

UNSET = -99.0  # "--tx/--ty/--tz not given" marker of the reference (:45-47,232-237)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Take a rgb encoded depth video and a color video, and view it/render as 3D")
    add = p.add_argument
    add("--depth_video", type=str, required=True, help="video file to use as input")
    add("--color_video", type=str, help="video file to use as color input")
    add("--xfov", type=int, help="fov in deg in the x-direction, calculated from aspect ratio and yfov if not given")
    add("--yfov", type=int, help="fov in deg in the y-direction, calculated from aspect ratio and xfov if not given")
    add("--max_depth", default=100, type=int, help="the max depth that the video uses")
    add("--render", action="store_true", help="Render to video instead of GUI")
    add("--render_as_pointcloud", action="store_true", help="Render as point cloud instead of as mesh")
    add("--remove_edges", action="store_true", help="Tries to remove edges that were not visible in the image")
    add("--show_camera", action="store_true", help="Shows lines representing the camera frustum")
    add("--background_ply", type=str, help="PLY file that will be included in the scene")
    add("--mask_video", type=str, help="Mask video to filter out back or foreground")
    add("--invert_mask", action="store_true", help="Remove the background (black) instead of the foreground (white)")
    add("--compressed", action="store_true", help="Render the video in a compressed format.")
    add("--draw_frame", default=-1, type=int, help="open gui with specific frame")
    add("--max_frames", default=-1, type=int, help="quit after max_frames nr of frames")
    add("--transformation_file", type=str, help="file with scene transformations from the aligner")
    add("--transformation_lock_frame", default=0, type=int, help="the frame that the transformation will use as a base")
    add("--x", default=2.0, type=float, help="camera x coordinate in meters")
    add("--y", default=2.0, type=float, help="camera y coordinate in meters")
    add("--z", default=-4.0, type=float, help="camera z coordinate in meters")
    add("--tx", default=-99.0, type=float, help="camera target x coordinate in meters")
    add("--ty", default=-99.0, type=float, help="camera target y coordinate in meters")
    add("--tz", default=-99.0, type=float, help="camera target z coordinate in meters")
    add("--chunk_frames", default=12, type=int, help="frames per GPU batch (addition; a multiple of 12 lets several decoders read one input)")
    add("--gpu_ffv1", action="store_true", help="code the FFV1 result video on the GPU (addition; also MDVT_FFV1_WRITER=gpu): the rendered "
        "frames never leave the device uncompressed; same container and codec, frames decode bit-identically")
    return p


def main(argv: Optional[List[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    if args.xfov is None and args.yfov is None:
        print("Either --xfov or --yfov is required.")
        return 0
    if not os.path.isfile(args.depth_video):
        raise Exception("input video does not exist")
    if args.color_video is not None and not os.path.isfile(args.color_video):
        raise Exception("input color_video does not exist")
    if not args.render or args.draw_frame != -1:
        raise NotImplementedError("the interactive Open3D viewer is a GUI feature; use --render")
    if args.mask_video:
        # In the reference the flag only works together with --remove_edges: without it create_mesh_from_point_cloud reads
        # `normals` before assignment (depth_map_tools.py:1346, UnboundLocalError); with it the result depends on the
        # per-triangle 89-degree test of the mesh, which the point splat does not build.
        raise NotImplementedError("--mask_video filters mesh triangles (and crashes the reference without --remove_edges); "
                                  "the point-splat render path has no triangles to filter")
    for flag in ("background_ply", "show_camera"):
        if getattr(args, flag):
            raise NotImplementedError(f"--{flag} adds scene objects the point-splat render path does not draw")
    transformations = None
    if args.transformation_file is not None:
        if not os.path.isfile(args.transformation_file):
            raise Exception("input transformation_file does not exist")
        with open(args.transformation_file) as fh:
            transformations = rebase_transformations(json.load(fh), args.transformation_lock_frame)

    w, h, fps, total = video_io.video_info(args.depth_video)
    fovx, fovy = fov_from_camera_matrix(compute_camera_matrix(args.xfov, args.yfov, w, h))
    print("Camera fovx: ", fovx, "fovy:", fovy)
    total_frames = total if args.max_frames < 0 else min(total, args.max_frames)
    device = torch.device("cuda", torch.cuda.current_device())
    target = tuple(None if v == UNSET else v for v in (args.tx, args.ty, args.tz))
    renderer = NovelViewRenderer(NovelViewParams(w, h, args.xfov, args.yfov, args.max_depth, (args.x, args.y, args.z), target, transformations,
                                                 of_by_one=not args.render_as_pointcloud), device)
    output_file, fourcc = (args.depth_video + "_render.mp4", "avc1") if args.compressed else (args.depth_video + "_render.mkv", "FFV1")
    lanes = video_io.default_lanes()
    on_device = fourcc == "FFV1" and video_io.gpu_ffv1_requested(getattr(args, "gpu_ffv1", False))
    if on_device:
        from .. import ffv1_gpu

        writer = ffv1_gpu.GpuFfv1Writer(output_file, fps, (w, h), device=device, batch=min(16, max(1, args.chunk_frames)))
    else:
        writer = video_io.ParallelWriter(output_file, fps, (w, h), lanes=lanes) if (fourcc == "FFV1" and lanes > 1) else \
            video_io.ChunkWriter(output_file, fourcc, fps, (w, h))
    done = 0
    host_out = None
    for n, (depth_rgb, colour) in video_io.open_chunk_reader([args.depth_video, args.color_video], 0, total_frames, chunk=args.chunk_frames,
                                                          decoders=video_io.default_decoders()):
        d = depth_rgb.to(device, non_blocking=True)
        c = d if colour is None else colour.to(device, non_blocking=True)
        rgb, _ = renderer.render_device(d, c, done)
        if on_device:   # the writer takes its own device copy: no raw frame crosses PCIe
            writer.write(rgb[:n])
            torch.cuda.current_stream(device).synchronize()   # this stream only (the renderer's second stream is joined to it): the writer's and the reader's streams keep running
            done += n
            print(f"Frame: {done} {done / fps}s", end="\r", file=sys.stderr)
            continue
        if host_out is None:
            host_out = torch.empty((args.chunk_frames, h, w, 3), dtype=torch.uint8, pin_memory=True)
        host_out[:n].copy_(rgb, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        writer.write(host_out[:n])
        done += n
        print(f"Frame: {done} {done / fps}s", end="\r", file=sys.stderr)
    writer.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
