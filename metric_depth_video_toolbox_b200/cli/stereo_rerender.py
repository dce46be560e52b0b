"""stereo_rerender.py front end: same flags, defaults, validation messages and output file names as the
reference script (stereo_rerender.py:270-968); the frame loop body runs on the GPU in chunks through
`StereoRerenderer` while OpenCV decodes / encodes on background threads.

    python stereo_rerender.py --depth_video D.mkv --color_video C.mkv --xfov 60 [--infill_mask ...]
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 stereo_rerender.py ...    # frames sharded over 8 GPUs

Deviations from the reference, all deliberate:
  * rendering is a point splat (the reference's own --render_as_pointcloud visibility rule), so
    --render_as_pointcloud / --remove_edges / --dont_remove_edges change nothing;
  * --max_frames N renders exactly N frames (the reference renders N+1 and then fails its own frame-count
    check, stereo_rerender.py:468,943,952);
  * under torchrun each rank renders a contiguous frame range into a segment file and rank 0 joins them.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from typing import List, Optional

import numpy as np
import torch

from .. import depth_frames_helper, sharding, video_io
from ..geometry import convergence_angle, curve_fit, fill_nan_with_closest, rebase_transformations  # noqa: F401
from ..depth_map_tools import timer  # noqa: F401  (stereo_rerender.py:15-20)
from ..infill import infill_using_normals, make_infill_mask, masked_blur  # noqa: F401  (module-level helpers other reference scripts import from stereo_rerender: basic_nomal_infill.py:10)
from ..stereo import StereoParams, StereoRerenderer
from ..vr180 import convert_to_equirectangular  # noqa: F401

UNSUPPORTED = {
    "mask_video": "--mask_video (background accumulation) is sequential host state outside the per-frame GPU path",
    "save_background": "--save_background belongs to the --mask_video background accumulation",
    "load_background": "--load_background belongs to the --mask_video background accumulation",
}


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Convert an RGB-encoded depth video and optional color video into a stereoscopic 3D "
                                            "side-by-side output for VR/3D TVs.")
    add = p.add_argument
    add("--master_xfov", type=float, default=45.0, help="Intended master FOV: how large the screen is from the viewer's point of view.")
    add("--depth_video", type=str, required=True, help="Path to input depth-encoded video file")
    add("--color_video", type=str, help="Path to input color video file")
    add("--xfov", type=float, help="Horizontal FOV in degrees")
    add("--yfov", type=float, help="Vertical FOV in degrees, calculated from aspect ratio and xfov if not given")
    add("--xfov_file", type=str, help="JSON file specifying xfov per frame")
    add("--max_depth", default=100, type=int, help="Maximum depth encoded in video")
    add("--transformation_file", type=str, help="file with scene transformations from the aligner")
    add("--transformation_lock_frame", default=0, type=int, help="the frame that the transformation will use as a base")
    add("--pupillary_distance", default=63, type=int, help="pupillary distance in mm")
    add("--max_frames", default=-1, type=int, help="Stop after processing this many frames")
    add("--touchly0", action="store_true", help="Render in touchly0 format (stereo 3D)")
    add("--vr180", action="store_true", help="Render in VR180 180 degree side-by-side")
    add("--render_as_pointcloud", action="store_true", help="Render output as point cloud (always the case here)")
    add("--convergence_file", type=str, help="json file with convergence data for each frame.")
    add("--dont_place_points_in_edges", action="store_true", help="Skip adding edge points for infill")
    add("--dont_remove_edges", action="store_true", help="Skip removing edges")
    add("--do_basic_infill", action="store_true", help="Use basic in-house infill algorithm.")
    add("--touchly1", action="store_true", help="Render in touchly1 format (mono+depth)")
    add("--touchly_max_depth", default=5, type=float, help="the max depth that touchly is clipped to.")
    add("--touchly_min_depth", default=0, type=float, help="the min depth that touchly is clipped to.")
    add("--compressed", action="store_true", help="Compress output video (lower quality)")
    add("--infill_mask", action="store_true", help="Save infill masks alongside output")
    add("--green_and_black_infill_mask", action="store_true", help="Dont generate normals for the infill mask.")
    add("--remove_edges", action="store_true", help="Remove mesh edges not visible in input frames")
    add("--mask_video", type=str, help="mask video used to build a background-only model")
    add("--save_background", action="store_true", help="Save the compound background as a file.")
    add("--load_background", help="Load the compound background from a file.")
    add("--create_sbs_depth_video", action="store_true", help="Save a depth version of the final sbs video")
    # additions (not in the reference)
    add("--chunk_frames", default=12, type=int, help="frames per GPU batch (a multiple of 12, OpenCV's FFV1 key-frame interval, lets several decoders work on one input)")
    add("--gpu_ffv1", action="store_true", help="code the FFV1 result videos on the GPU (addition; also MDVT_FFV1_WRITER=gpu): same "
        "container and codec, every frame a key frame and ~1000 slices per frame, frames decode bit-identically")
    add("--writer_lanes", default=0, type=int, help="parallel FFV1 encoder lanes per output file (0: from the host core count, 1: the reference's single writer)")
    return p


def _require_file(path: Optional[str], what: str, exc=FileNotFoundError):
    if path and not os.path.isfile(path):
        raise exc(f"{what} not found: {path}")


def load_clip_parameters(args, frame_width: int, frame_height: int, n_frames: int, rendered_frames: Optional[int] = None) -> StereoParams:
    """Everything of stereo_rerender.py:343-373,402-404 that spans the whole clip (rank 0 only under torchrun).
    n_frames: frames of the clip; rendered_frames: frames this job renders (--max_frames): the per-frame lists of a
    convergence / transformation file only have to cover those, as in the reference, which indexes them by frame."""
    rendered_frames = n_frames if rendered_frames is None else rendered_frames
    convergence = None
    if args.convergence_file:
        _require_file(args.convergence_file, "Convergence file")
        with open(args.convergence_file) as fh:
            convergence = curve_fit(fill_nan_with_closest(json.load(fh)))
    xfovs = None
    if args.xfov_file:
        _require_file(args.xfov_file, "XFOV file")
        with open(args.xfov_file) as fh:
            xfovs = json.load(fh)
        if not isinstance(xfovs, list) or not all(isinstance(x, (int, float)) for x in xfovs):
            raise ValueError("XFOV file must contain a list of numbers.")
        if len(xfovs) != n_frames:
            raise ValueError(f"XFOV file must have the same number of frames as the input video ({n_frames} vs xfov={len(xfovs)}).")
    transformations = None
    if args.transformation_file is not None:
        if not os.path.isfile(args.transformation_file):
            raise Exception("input transformation_file does not exist")
        with open(args.transformation_file) as fh:
            transformations = rebase_transformations(json.load(fh), args.transformation_lock_frame)
    for name, seq in (("convergence", convergence), ("transformation", transformations)):
        if seq is not None and len(seq) < rendered_frames:
            raise ValueError(f"{name} file has {len(seq)} entries, the job renders {rendered_frames} frames")
    def whole_clip(seq):  # per-frame list cut / padded (last entry repeated; never rendered) to the clip's length: fixed broadcast format
        seq = list(seq)[:n_frames]
        return seq + [seq[-1]] * (n_frames - len(seq))

    return StereoParams(frame_width, frame_height, xfov=args.xfov, yfov=args.yfov, xfovs=xfovs, max_depth=args.max_depth,
                        pupillary_distance=args.pupillary_distance, master_xfov=args.master_xfov,
                        convergence_depths=None if convergence is None else whole_clip(convergence),
                        transformations=None if transformations is None else whole_clip(transformations),
                        infill_mask=bool(args.infill_mask), mask_rgb=True)


def output_names(args):
    """stereo_rerender.py:409-435."""
    kind = "Touchly1" if args.touchly1 else ("Touchly0" if args.touchly0 else "stereo")
    ext, fourcc = ("mp4", "avc1") if args.compressed else ("mkv", "FFV1")
    return f"{args.depth_video}_{kind}.{ext}", f"{args.depth_video}_tmp_{kind}.{ext}", fourcc


def main(argv: Optional[List[str]] = None) -> int:
    return run(build_parser().parse_args(argv))


def run(args, keep_process_group: bool = False) -> int:
    """The script body for already-parsed arguments.  keep_process_group: leave torch.distributed initialised (a caller
    that renders several clips in one torchrun job, e.g. movie_steps.step5_render_sbs)."""
    if args.xfov is None and args.yfov is None and args.xfov_file is None:
        raise ValueError("Error: Either --xfov_file, --xfov or --yfov must be provided.")
    if args.green_and_black_infill_mask and args.do_basic_infill:
        raise ValueError("Error: --green_and_black_infill_mask and --do_basic_infill are not compatible with eachother.")
    for flag, why in UNSUPPORTED.items():
        if getattr(args, flag):
            raise NotImplementedError(why)
    _require_file(args.depth_video, "Depth video")
    _require_file(args.color_video, "Color video")

    rank, world_size, local_rank = sharding.init_from_env()
    device = torch.device("cuda", local_rank)

    frame_width, frame_height, frame_rate, total = video_io.video_info(args.depth_video)
    if args.color_video:
        cw, ch, cfps, _ = video_io.video_info(args.color_video)
        if (frame_width, frame_height) != (cw, ch):
            raise ValueError(f"Depth video and Color video must have the same dimensions (Depth: {frame_width}x{frame_height} vs Color {cw}x{ch}).")
        if round(frame_rate, 2) != round(cfps, 2):
            raise ValueError(f"Color video and depth video must have the same frame rate (Depth={frame_rate} vs Color={round(cfps, 2)}).")
    total_frames = total if args.max_frames < 0 else min(total, args.max_frames)

    # whole-clip parameters: built once on rank 0, broadcast (the only inter-GPU traffic of the job)
    params = None
    if rank == 0:
        try:
            params = load_clip_parameters(args, frame_width, frame_height, total, total_frames)
        except Exception as exc:  # travels to the other ranks as a failure marker: nobody is left waiting in a broadcast
            params = exc
    params, _ = sharding.broadcast_params(params, total)
    # GOP-aligned ranges when every rank still gets work: range starts are key frames of the input files
    align = video_io.GOP if total_frames >= world_size * video_io.GOP else 1
    start, stop = sharding.frame_range(total_frames, rank, world_size, align)

    from . import stereo_modes

    job = stereo_modes.StereoJob(args, params, device, frame_width, frame_height)
    output_file, output_tmp_file, fourcc = output_names(args)
    seg = (lambda p: p) if world_size == 1 else (lambda p: f"{p}.rank{rank:02d}.mkv")
    # FFV1 results go through parallel encoder lanes joined at packet level (video_io.ParallelWriter); the compressed
    # (avc1) variant keeps the single writer and, under torchrun, the frame-copy join
    lanes = args.writer_lanes if args.writer_lanes > 0 else video_io.default_lanes(world_size)
    parallel = fourcc == "FFV1" and lanes > 1

    gpu_ffv1 = video_io.gpu_ffv1_requested(getattr(args, "gpu_ffv1", False))
    job.device_outputs = gpu_ffv1 and fourcc == "FFV1"   # plain stereo mode: SBS + mask go to the coder without a host round trip

    def open_writer(path: str, cc: str):
        if gpu_ffv1 and cc == "FFV1":   # entropy coding on the device; ranks leave segments + plans for the packet-level join
            from .. import ffv1_gpu

            return ffv1_gpu.GpuFfv1Writer(seg(path), frame_rate, job.out_size, device=device, batch=min(16, max(1, args.chunk_frames)),
                                          join_on_close=(world_size == 1))
        if parallel and cc == "FFV1":
            return video_io.ParallelWriter(seg(path), frame_rate, job.out_size, lanes=lanes, join_on_close=(world_size == 1))
        return video_io.ChunkWriter(seg(path), cc if world_size == 1 else "FFV1", frame_rate, job.out_size)

    writers = {"main": open_writer(output_tmp_file, fourcc)}
    if job.writes_mask:
        writers["mask"] = open_writer(output_tmp_file + "_infillmask.mkv", "FFV1")
    if job.has_depth_output:
        writers["depth"] = open_writer(output_tmp_file + "_depth.mkv", "FFV1")

    reader = video_io.open_chunk_reader([args.depth_video, args.color_video], start, stop, chunk=args.chunk_frames, device=device,
                                  decoders=video_io.default_decoders(world_size))
    # The loop shares the interpreter with the reader's and the writers' threads (packet copies, muxing).  With CPython's
    # default 5 ms switch interval every GIL hand-back costs this thread up to 5 ms -- measured: 4-7 ms of "render" per frame
    # for 0.2 ms of kernels -- so the interval is shortened for the duration of the job.
    switch_interval = sys.getswitchinterval()
    sys.setswitchinterval(1e-4)
    t0 = time.time()
    done = 0
    deferred = {}
    stage = {"read": 0.0, "render": 0.0, "write": 0.0}   # MDVT_PROFILE_LOOP=1: where the frame loop's wall time goes
    t_prev = time.perf_counter()
    for n, (depth_rgb, colour) in reader:
        t_a = time.perf_counter()
        outputs = job.render_chunk(depth_rgb, depth_rgb if colour is None else colour, start + done)
        # results are complete and the reader may recycle its buffers once THIS stream is done (the renderer's second stream
        # is joined to it by an event); a device-wide synchronize would also wait for the writers' encodes and the reader's
        # decodes of other chunks, which run on streams of their own precisely so that they overlap this loop
        torch.cuda.current_stream(device).synchronize()
        t_b = time.perf_counter()
        for key, pending in list(deferred.items()):   # last chunk's host-finished frames (TELEA tail of the mask)
            writers[key].write(pending.result(), rgb=(key != "depth"))
            del deferred[key]
        for key, w in writers.items():
            if hasattr(outputs[key], "result"):
                deferred[key] = outputs[key]
            else:
                w.write(outputs[key], rgb=(key != "depth"))
        done += n
        t_c = time.perf_counter()
        stage["read"] += t_a - t_prev
        stage["render"] += t_b - t_a
        stage["write"] += t_c - t_b
        t_prev = t_c
        if rank == 0:
            pct = 100.0 * done / max(1, stop - start)
            print(f"[{pct:5.1f}%] Frame #{done:4d}/{stop - start}  {done / max(1e-9, time.time() - t0):7.1f} frames/s", end="\r", file=sys.stderr)
    for key, pending in deferred.items():
        writers[key].write(pending.result(), rgb=(key != "depth"))
    t_close = time.perf_counter()
    for w in writers.values():
        w.close()
    sys.setswitchinterval(switch_interval)
    if os.environ.get("MDVT_PROFILE_LOOP") == "1":
        print(f"\n[frame loop, rank {rank}] {done} frames: waiting for the reader {stage['read']:.2f} s, render + sync {stage['render']:.2f} s, "
              f"handing to the writers {stage['write']:.2f} s, closing the writers {time.perf_counter() - t_close:.2f} s "
              f"({type(reader).__name__}, {', '.join(type(w).__name__ for w in writers.values())})", file=sys.stderr)
    total_done = sharding.gather_counts(done)

    if world_size > 1:
        import torch.distributed as dist

        dist.barrier()
        if rank == 0:
            for suffix, cc in (("", fourcc), ("_infillmask.mkv", "FFV1"), ("_depth.mkv", "FFV1")):
                parts = [f"{output_tmp_file}{suffix}.rank{r:02d}.mkv" for r in range(world_size)]
                if all(os.path.isfile(p + ".plan.json") for p in parts):   # parallel lanes: packet-level join, no re-encode
                    video_io.join_plans([video_io.load_plan(p) for p in parts], output_tmp_file + suffix, frame_rate)
                elif all(os.path.isfile(p) for p in parts):
                    stereo_modes.join_segments(parts, output_tmp_file + suffix, cc, frame_rate, job.out_size)
        dist.barrier()
    if rank == 0:
        depth_frames_helper.verify_and_move(output_tmp_file, total_frames, output_file)
        if "mask" in writers:
            depth_frames_helper.verify_and_move(output_tmp_file + "_infillmask.mkv", total_frames, output_file + "_infillmask.mkv")
        if "depth" in writers:
            depth_frames_helper.verify_and_move(output_tmp_file + "_depth.mkv", total_frames, output_file + "_depth.mkv")
        print(f"\nProcessing complete ({total_done} frames, {total_done / max(1e-9, time.time() - t0):.1f} frames/s). Output saved to: {output_file}")
    if world_size > 1 and not keep_process_group:
        import torch.distributed as dist

        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
