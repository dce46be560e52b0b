"""Novel-view render of a clip: the body of `3d_view_depthfile.py --render`'s frame loop (:133-255) for
batches of frames.  Per frame: decode -> unproject -> optional pose -> look-at camera aimed at the vertex
centroid (Open3D get_center(), :231) -> z-buffered splat -> white-background image.

The centroid is a per-frame reduction that decides the camera of the same frame.  `render_device` runs the whole
chunk in ONE library call with no host round trip (`mdvt_novel_view_frames`): per frame the reduction, a
finishing kernel that also evaluates the look-at camera in float64 on the device, the splat reading that camera
from device memory, and the resolve.  `render_device_hostcam` is the two-pass form (all centroids of the chunk,
one small D2H copy, NumPy look-at matrices, splat + resolve per frame) kept as the cross-check of the device
camera."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import geometry as geo
from . import ops


@dataclass
class NovelViewParams:
    """Mirrors the 3d_view_depthfile.py CLI values that reach the per-pixel path (:23-47)."""
    width: int
    height: int
    xfov: Optional[float] = None
    yfov: Optional[float] = None
    max_depth: float = 100
    cam_pos: Sequence[float] = (2.0, 2.0, -4.0)              # --x --y --z
    target: Sequence[Optional[float]] = (None, None, None)   # --tx --ty --tz (None: centroid component)
    transformations: Optional[Sequence] = None               # per-frame 4x4, already re-based
    of_by_one: bool = True        # vertex grid used for the centroid (mesh mode; False with --render_as_pointcloud, :178-180)
    bg_rgb: Sequence[int] = (255, 255, 255)                  # :254
    near: float = geo.NEAR_PLANE


class NovelViewRenderer:
    def __init__(self, params: NovelViewParams, device: Optional[torch.device] = None):
        self.p = params
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.K = geo.compute_camera_matrix(params.xfov, params.yfov, params.width, params.height)
        self._zbuf = None
        self._sums = None
        self._views = None
        self._csums = None

    def _pose(self, frame: int):
        return None if self.p.transformations is None else np.asarray(self.p.transformations[frame], dtype=np.float64)

    def centroids(self, depth_rgb: torch.Tensor, start_frame: int = 0) -> np.ndarray:
        """(n, 3) float64 vertex means of the chunk (host)."""
        p = self.p
        n = depth_rgb.shape[0]
        stride = 4 + _lib.REDUCE_SCRATCH_DOUBLES
        if self._csums is None or self._csums.shape[0] < n or self._csums.device != depth_rgb.device:
            self._csums = torch.empty((n, stride), dtype=torch.float64, device=depth_rgb.device)
        src = ops.make_source(p.width, p.height, self.K, p.max_depth, "D1", True, 1.0, p.of_by_one)
        for k in range(n):
            ops.centroid_sums(depth_rgb[k], src, self.K, self._pose(start_frame + k), out=self._csums[k])
        s = self._csums[:n, :4].cpu().numpy()  # the one synchronising copy of the chunk
        return s[:, :3] / s[:, 3:4]

    def extrinsic(self, centroid: np.ndarray) -> np.ndarray:
        look = np.array(centroid, dtype=np.float64)
        for a in range(3):
            if self.p.target[a] is not None:
                look[a] = self.p.target[a]
        return geo.cam_look_at(np.array(self.p.cam_pos).astype(np.float32), look)

    def view_of(self, frame: int, centroid: np.ndarray) -> ops.ViewSpec:
        """render()'s camera: geometry Y scaled by fy/fx before the extrinsic, fx on both axes
        (depth_map_tools.py:1528-1552); Open3D uses the upper 3x4 of the look-at matrix."""
        K = self.K
        M = self.extrinsic(centroid)[:3, :4] @ np.diag([1.0, K[1, 1] / K[0, 0], 1.0, 1.0])
        pose = self._pose(frame)
        if pose is not None:
            M = M @ pose
        return ops.ViewSpec(M, K[0, 0], K[0, 0], K[0, 2], K[1, 2])

    def _outputs(self, depth_rgb, out_rgb, out_mask):
        p = self.p
        n, h, w, _ = depth_rgb.shape
        if (w, h) != (p.width, p.height):
            raise ValueError(f"frames are {w}x{h}, parameters say {p.width}x{p.height}")
        dev = depth_rgb.device
        if out_rgb is None:
            out_rgb = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
        if out_mask is None:
            out_mask = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
        sets = 1 if os.environ.get("MDVT_ZBUF_SETS", "2") == "1" else 2   # two planes: frames alternate between two streams
        if self._zbuf is None or tuple(self._zbuf.shape) != (sets, h, w) or self._zbuf.device != dev:
            self._zbuf = ops.new_zbuf(sets, w, h, dev)
        return n, h, w, out_rgb, out_mask

    def render_device(self, depth_rgb: torch.Tensor, colour: torch.Tensor, start_frame: int = 0, out_rgb: Optional[torch.Tensor] = None,
                      out_mask: Optional[torch.Tensor] = None):
        """depth_rgb / colour (n, H, W, 3) u8 CUDA -> (rgb (n, H, W, 3) u8, hole mask (n, H, W) u8).  Asynchronous: one
        library call, cameras computed on the device (`self.last_views`: (n, 16) float32, `self.last_sums`)."""
        p = self.p
        n, h, w, out_rgb, out_mask = self._outputs(depth_rgb, out_rgb, out_mask)
        stride = 4 + _lib.REDUCE_SCRATCH_DOUBLES
        if self._sums is None or self._sums.shape[0] < n or self._sums.device != depth_rgb.device:
            self._sums = torch.empty((n, stride), dtype=torch.float64, device=depth_rgb.device)
            self._views = torch.empty((n, 16), dtype=torch.float32, device=depth_rgb.device)
            self._touched = torch.empty(2 * int(_lib.load().mdvt_touched_bytes(w, h)), dtype=torch.uint8, device=depth_rgb.device)
        src_c = ops.make_source(w, h, self.K, p.max_depth, "D1", True, 1.0, p.of_by_one)
        src = ops.make_source(w, h, self.K, p.max_depth, "D1", True, 1.0, False)
        poses = None if p.transformations is None else np.stack([self._pose(start_frame + k) for k in range(n)])
        ops.novel_view_frames(depth_rgb, colour, src_c, src, self.K, p.cam_pos, p.target, poses, self._zbuf, out_rgb, out_mask, p.bg_rgb,
                              p.bg_rgb, 0, p.near, self._sums[:n], self._views[:n], self._touched)
        self.last_sums, self.last_views = self._sums[:n], self._views[:n]
        return out_rgb, out_mask

    def render_device_hostcam(self, depth_rgb: torch.Tensor, colour: torch.Tensor, start_frame: int = 0,
                              out_rgb: Optional[torch.Tensor] = None, out_mask: Optional[torch.Tensor] = None):
        """Same result with the cameras evaluated on the host (synchronises once per chunk)."""
        p = self.p
        n, h, w, out_rgb, out_mask = self._outputs(depth_rgb, out_rgb, out_mask)
        centres = self.centroids(depth_rgb, start_frame)
        src = ops.make_source(w, h, self.K, p.max_depth, "D1", True, 1.0, False)
        views = [[self.view_of(start_frame + k, centres[k])] for k in range(n)]
        ops.render_views(depth_rgb, colour, [src], views, w, h, self._zbuf[:1], out_rgb, out_mask, None, p.bg_rgb, p.bg_rgb, 0, p.near)
        return out_rgb, out_mask

    def render_host(self, depth_rgb, colour, out_rgb=None, start_frame: int = 0, chunk_frames: int = 2):
        """Host arrays in ((n, H, W, 3) u8, pinned memory makes the copies asynchronous), pinned host tensor out; the returned
        buffer is complete.  Three streams: the uploads of chunk i + 1 and the download of chunk i - 1 run next to the kernels of
        chunk i (the kernels themselves stay on the caller's stream: the frame loop keeps per-renderer state and an internal
        second stream), through two sets of device staging buffers of `chunk_frames` frames."""
        d_host, c_host = torch.as_tensor(depth_rgb), torch.as_tensor(colour)
        n, h, w, _ = d_host.shape
        if out_rgb is None:
            out_rgb = torch.empty((n, h, w, 3), dtype=torch.uint8, pin_memory=True)
        out_t = torch.as_tensor(out_rgb)
        chunk = max(1, min(int(os.environ.get("MDVT_HOST_CHUNK", chunk_frames)), n))
        dev = self.device
        key = (chunk, h, w)
        if getattr(self, "_host_key", None) != key:
            self._host = dict(up=torch.cuda.Stream(device=dev), down=torch.cuda.Stream(device=dev),
                              slots=[dict(d=torch.empty((chunk, h, w, 3), dtype=torch.uint8, device=dev),
                                          c=torch.empty((chunk, h, w, 3), dtype=torch.uint8, device=dev),
                                          rgb=torch.empty((chunk, h, w, 3), dtype=torch.uint8, device=dev),
                                          mask=torch.empty((chunk, h, w), dtype=torch.uint8, device=dev)) for _ in range(2)])
            self._host_key = key
        hp = self._host
        main = torch.cuda.current_stream(dev)
        hp["up"].wait_stream(main)
        hp["down"].wait_stream(main)
        starts = list(range(0, n, chunk))
        uploaded = [torch.cuda.Event() for _ in starts]
        rendered = [torch.cuda.Event() for _ in starts]
        downloaded = [torch.cuda.Event() for _ in starts]

        def upload(i):
            f0 = starts[i]
            cnt = min(chunk, n - f0)
            slot = hp["slots"][i % 2]
            with torch.cuda.stream(hp["up"]):
                if i >= 2:
                    hp["up"].wait_event(rendered[i - 2])   # the kernels that read this slot's inputs are done
                slot["d"][:cnt].copy_(d_host[f0:f0 + cnt], non_blocking=True)
                slot["c"][:cnt].copy_(c_host[f0:f0 + cnt], non_blocking=True)
                uploaded[i].record(hp["up"])

        upload(0)
        for i, f0 in enumerate(starts):
            cnt = min(chunk, n - f0)
            slot = hp["slots"][i % 2]
            if i + 1 < len(starts):
                upload(i + 1)
            main.wait_event(uploaded[i])
            if i >= 2:
                main.wait_event(downloaded[i - 2])        # this slot's previous result has left the device
            self.render_device(slot["d"][:cnt], slot["c"][:cnt], start_frame + f0, slot["rgb"][:cnt], slot["mask"][:cnt])
            rendered[i].record(main)
            with torch.cuda.stream(hp["down"]):
                hp["down"].wait_event(rendered[i])
                out_t[f0:f0 + cnt].copy_(slot["rgb"][:cnt], non_blocking=True)
                downloaded[i].record(hp["down"])
        main.wait_stream(hp["down"])
        hp["down"].synchronize()
        return out_rgb
