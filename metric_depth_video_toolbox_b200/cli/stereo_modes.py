"""Output modes of the stereo_rerender front end: what is rendered per chunk and how the output frame is
assembled (stereo_rerender.py:406-422,548-552,677-702,823-829,910-941)."""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from .. import video_io
from ..stereo import StereoParams, StereoRerenderer


class StereoJob:
    """One clip shard -> chunks of output frames on the host.

    render_chunk(depth_rgb, colour, first_frame) takes pinned host tensors (n, H, W, 3) u8 RGB and returns
    {"main": (n, oh, ow, 3) u8 RGB, "mask": ... (with --infill_mask), "depth": ... BGR (with
    --create_sbs_depth_video)} host tensors that stay valid until the next call."""

    def __init__(self, args, params: StereoParams, device: torch.device, frame_width: int, frame_height: int):
        self.args, self.params, self.device = args, params, device
        self.w, self.h = frame_width, frame_height
        for flag in ("touchly0", "touchly1", "vr180", "create_sbs_depth_video", "do_basic_infill"):
            if getattr(args, flag, False):
                raise NotImplementedError(f"--{flag} is not built yet in this front end")
        if args.infill_mask and not args.green_and_black_infill_mask:
            raise NotImplementedError("the normals-coded infill mask is not built yet: add --green_and_black_infill_mask")
        self.out_size = (2 * self.w, self.h)
        self.has_depth_output = False
        self.renderer = StereoRerenderer(params, device)
        self._host: Dict[str, torch.Tensor] = {}

    def _host_buf(self, key: str, shape) -> torch.Tensor:
        buf = self._host.get(key)
        if buf is None or buf.shape[1:] != tuple(shape[1:]) or buf.shape[0] < shape[0]:
            buf = self._host[key] = torch.empty(tuple(shape), dtype=torch.uint8, pin_memory=True)
        return buf[:shape[0]]

    def render_chunk(self, depth_rgb: torch.Tensor, colour: torch.Tensor, first_frame: int) -> Dict[str, torch.Tensor]:
        n = depth_rgb.shape[0]
        sbs = self._host_buf("main", (n, self.h, 2 * self.w, 3))
        mask = self._host_buf("mask", (n, self.h, 2 * self.w, 3)) if self.params.infill_mask else None
        self.renderer.render_host(depth_rgb, colour, sbs, mask, start_frame=first_frame, chunk_frames=max(1, min(n, 8)))
        out = {"main": sbs}
        if mask is not None:
            out["mask"] = mask
        return out


def join_segments(parts: List[str], out_path: str, fourcc: str, fps: float, size):
    """Concatenate per-rank segment files into one video (frame copy through OpenCV; FFV1 is lossless, so
    the joined file holds exactly the frames the ranks rendered) and delete the segments."""
    import os

    import cv2

    writer = cv2.VideoWriter(out_path, cv2.VideoWriter_fourcc(*fourcc), fps, size)
    for p in parts:
        cap = cv2.VideoCapture(p)
        while True:
            ok, frame = cap.read()
            if not ok:
                break
            writer.write(frame)
        cap.release()
    writer.release()
    for p in parts:
        os.remove(p)
