"""find_convergence_depth.py front end (reference :15-94): per-frame (masked) mean depth ->
`<depth_video>_convergence_depths.json`.  Decode (D3) + mask test + sum run in one reduction kernel per
frame; a chunk's sums come back in one small copy."""
from __future__ import annotations

import argparse
import json
import os
import sys
from typing import List, Optional

import torch

from .. import _lib, ops, video_io


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="finds convergence depth in depth video: the depth at which a video's main focus lies.")
    p.add_argument("--depth_video", type=str, required=True, help="Depth video file to analyse")
    p.add_argument("--mask_video", type=str, help="black and white mask video for the main focus area (white = area of interest)")
    p.add_argument("--max_depth", default=100, type=int, help="the max depth that the video uses")
    return p


def convergence_depths(depth_video: str, mask_video: Optional[str], max_depth, device, chunk: int = 12) -> List[float]:
    """find_convergence_depth.py:46-80: one value per frame of the DEPTH video.  Where the mask video has run out (or
    there is none) the mean is taken over the whole frame (:63-74, `mesured_pixels = depth`); frames whose mask selects
    no pixel give NaN.  The mean is the float64 sum / count rounded to float32 (the reference's float32 `.mean()`
    agrees with it to float32 rounding)."""
    import numpy as np

    out: List[float] = []
    sums = torch.empty((chunk, 4 + _lib.REDUCE_SCRATCH_DOUBLES), dtype=torch.float64, device=device)

    def analyse(paths, start):
        for n, (depth_rgb, mask) in video_io.open_chunk_reader(paths, start, None, chunk=chunk, grey=[False, True], decoders=video_io.default_decoders(), device=device):
            d = depth_rgb.to(device, non_blocking=True)
            m = None if mask is None else mask.to(device, non_blocking=True)
            for k in range(n):
                ops.depth_sums(d[k], max_depth, "D3", True, None if m is None else m[k], 240, out=sums[k])
            host = sums[:n, :2].cpu()
            for s, cnt in host.tolist():
                out.append(float(np.float32(s / cnt)) if cnt > 0 else float("nan"))

    analyse([depth_video, mask_video], 0)
    if mask_video is not None:  # the mask video ended first: the remaining frames are measured whole ("Failed to read mask video frame")
        analyse([depth_video, None], len(out))
    return out


def main(argv: Optional[List[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    if not os.path.isfile(args.depth_video):
        raise Exception("input color_video does not exist")
    if args.mask_video is not None and not os.path.isfile(args.mask_video):
        raise Exception("input mask_video does not exist")
    device = torch.device("cuda", torch.cuda.current_device())
    depths = convergence_depths(args.depth_video, args.mask_video, args.max_depth, device)
    with open(args.depth_video + "_convergence_depths.json", "w") as fh:
        fh.write(json.dumps(depths))
    return 0


if __name__ == "__main__":
    sys.exit(main())
