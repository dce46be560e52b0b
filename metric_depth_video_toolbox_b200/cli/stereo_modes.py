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

    render_chunk(depth_rgb, colour, first_frame) takes (n, H, W, 3) u8 RGB tensors -- pinned host chunks of video_io.ChunkReader
    or device chunks of video_io.DeviceChunkReader -- and returns {"main": (n, oh, ow, 3) u8 RGB, "mask": ... RGB (with
    --infill_mask), "depth": ... BGR wire format (with --create_sbs_depth_video)}: host tensors, or with `device_outputs` (the
    result videos are coded on the device) device tensors wherever no host pass remains; valid until the next call.

    Modes (stereo_rerender.py:406-422):
      stereo    left | right side by side (:910-918), green/black hole mask (:787-793,921-928), optional SBS depth
                video (:930-939)
      touchly1  colour over reverse depth (:548-552 without a pose file -- no render at all; :677-692 with one)
    """

    def __init__(self, args, params: StereoParams, device: torch.device, frame_width: int, frame_height: int):
        self.args, self.params, self.device = args, params, device
        self.w, self.h = frame_width, frame_height
        self.touchly0 = bool(getattr(args, "touchly0", False))
        self.vr180 = bool(getattr(args, "vr180", False)) or self.touchly0       # :406-407
        if self.vr180 and args.infill_mask:
            raise NotImplementedError("--vr180 / --touchly0 with --infill_mask: the reference itself fails here (a 1920x1920 hole "
                                      "mask indexed into a frame-sized mask image, stereo_rerender.py:527-528,740,787-792)")
        self.basic_infill = bool(getattr(args, "do_basic_infill", False))
        if self.basic_infill and not args.infill_mask:
            raise NotImplementedError("--do_basic_infill without --infill_mask marks holes with a black background, which a black pixel "
                                      "of the film is indistinguishable from; pass --infill_mask as movie_2_3D does")
        # stereo_rerender.py:568-573,589: the mesh edge test runs for --infill_mask / --remove_edges / --do_basic_infill unless
        # --dont_remove_edges; its vertices are painted into the holes unless --dont_place_points_in_edges
        remove_edges = (args.infill_mask or args.remove_edges or self.basic_infill) and not args.dont_remove_edges
        self.paint_edges = bool(remove_edges and not args.dont_place_points_in_edges)
        self.code_normals = bool(args.infill_mask and not args.green_and_black_infill_mask)
        self.touchly1 = bool(getattr(args, "touchly1", False))
        if self.touchly1 and params.transformations is not None and args.infill_mask:
            raise NotImplementedError("--touchly1 with a pose file and --infill_mask: the reference itself fails here "
                                      "(cv2.cvtColor(RGB2BGR) on its single-channel mask, stereo_rerender.py:701-702)")
        self.out_size = (self.w, 2 * self.h) if self.touchly1 else (2 * self.w, self.h)
        if self.vr180 and not self.touchly1:
            self.vr_side = 1920                                                 # out_width, out_height = 1920, 1920 (:528)
            self.out_size = ((3 if self.touchly0 else 2) * self.vr_side, self.vr_side)
        self.has_depth_output = bool(getattr(args, "create_sbs_depth_video", False)) and not self.touchly1
        self.writes_mask = bool(args.infill_mask) and not self.touchly1  # the touchly1 fast path never writes mask frames
        self.renderer = StereoRerenderer(params, device)
        self.infill = None
        if (self.paint_edges or self.code_normals) and not self.touchly1 and not self.vr180:
            from ..infill import InfillMaskRenderer

            self.infill = InfillMaskRenderer(self.renderer)
        self._host: Dict[str, torch.Tensor] = {}
        self._dev: Dict[str, torch.Tensor] = {}
        self._zbuf = None
        # True (set by the front end when the result videos are coded on the device): the plain stereo mode hands out its
        # device tensors instead of downloading them -- valid until the next render_chunk, the writer takes its own copy
        self.device_outputs = False

    def _host_buf(self, key: str, shape) -> torch.Tensor:
        buf = self._host.get(key)
        if buf is None or buf.shape[1:] != tuple(shape[1:]) or buf.shape[0] < shape[0]:
            buf = self._host[key] = torch.empty(tuple(shape), dtype=torch.uint8, pin_memory=True)
        return buf[:shape[0]]

    def _dev_buf(self, key: str, shape, dtype=torch.uint8) -> torch.Tensor:
        buf = self._dev.get(key)
        if buf is None or buf.shape[1:] != tuple(shape[1:]) or buf.shape[0] < shape[0] or buf.dtype != dtype:
            buf = self._dev[key] = torch.empty(tuple(shape), dtype=dtype, device=self.device)
        return buf[:shape[0]]

    # ---- stereo ---------------------------------------------------------------------------------------
    def _stereo_chunk(self, depth_rgb, colour, first_frame):
        n = depth_rgb.shape[0]
        deferred_mask = None
        sbs = self._host_buf("main", (n, self.h, 2 * self.w, 3))
        mask = self._host_buf("mask", (n, self.h, 2 * self.w, 3)) if self.params.infill_mask else None
        if self.device_outputs and not self.has_depth_output and self.infill is None:
            d = self._dev_buf("d", depth_rgb.shape)
            c = self._dev_buf("c", colour.shape)
            d.copy_(depth_rgb, non_blocking=True)
            c.copy_(colour, non_blocking=True)
            dsbs = self._dev_buf("sbs", (n, self.h, 2 * self.w, 3))
            dmask = self._dev_buf("mask", (n, self.h, 2 * self.w, 3)) if self.params.infill_mask else None
            self.renderer.render_device(d, c, first_frame, dsbs, dmask, None)
            out = {"main": dsbs}
            if dmask is not None:
                out["mask"] = dmask
            return out
        if not self.has_depth_output and self.infill is None:  # the pipelined two-stream path
            self.renderer.render_host(depth_rgb, colour, sbs, mask, start_frame=first_frame, chunk_frames=max(1, min(n, 4)))
            out = {"main": sbs}
        elif self.infill is not None:  # edge points painted into the holes, normals-coded (or green/black) mask image
            from .. import ops

            d = self._dev_buf("d", depth_rgb.shape)
            c = self._dev_buf("c", colour.shape)
            d.copy_(depth_rgb, non_blocking=True)
            c.copy_(colour, non_blocking=True)
            dsbs = self._dev_buf("sbs", (n, self.h, 2 * self.w, 3))
            dmask = self._dev_buf("mask", (n, self.h, 2 * self.w, 3))
            ddepth = self._dev_buf("depth", (n, self.h, 2 * self.w), torch.float32) if self.has_depth_output else None
            # with --do_basic_infill the holes are filled by the normal march instead of the edge colours (:810-814); the
            # edge points still contribute their normals to the mask
            self.infill.render_device(d, c, first_frame, dsbs, dmask, self.code_normals, self.paint_edges and not self.basic_infill, ddepth)
            if self.device_outputs and not self.code_normals:
                # nothing left to do on the host (no TELEA pass): the frames go to the device coder as they are
                out = {"main": dsbs}
                if ddepth is not None:
                    out["depth"] = ops.encode_depth(ddepth, self.params.max_depth, True, True)
                if mask is not None:
                    out["mask"] = dmask
                return out
            if not self.basic_infill:
                sbs.copy_(dsbs, non_blocking=True)
            out = {"main": sbs}
            if ddepth is not None:
                coded = ops.encode_depth(ddepth, self.params.max_depth, True, True)
                hdepth = self._host_buf("depthcode", coded.shape)
                hdepth.copy_(coded, non_blocking=True)
                out["depth"] = hdepth
            if mask is not None:
                mask.copy_(dmask, non_blocking=True)
                if self.code_normals:  # host part: TELEA + masked blur per eye (OpenCV, as the reference)
                    torch.cuda.synchronize(self.device)
                    if self.basic_infill:  # final mask back to the device, march, filled image to the host
                        mask.copy_(torch.from_numpy(self.infill.finish(mask.numpy())))
                        dmask.copy_(mask, non_blocking=True)
                        self.infill.basic_infill(dsbs, dmask)
                        sbs.copy_(dsbs, non_blocking=True)
                    else:  # finished on the worker pool while the loop goes on; the front end writes it one chunk later
                        deferred_mask = self.infill.finish_async(mask.numpy())
        else:
            from .. import ops

            d = self._dev_buf("d", depth_rgb.shape)
            c = self._dev_buf("c", colour.shape)
            d.copy_(depth_rgb, non_blocking=True)
            c.copy_(colour, non_blocking=True)
            dsbs = self._dev_buf("sbs", (n, self.h, 2 * self.w, 3))
            dmask = self._dev_buf("mask", (n, self.h, 2 * self.w, 3)) if mask is not None else None
            ddepth = self._dev_buf("depth", (n, self.h, 2 * self.w), torch.float32)
            self.renderer.render_device(d, c, first_frame, dsbs, dmask, ddepth)
            coded = ops.encode_depth(ddepth, self.params.max_depth, True, True)  # B, G, R like encode_data_as_BGR (:932-936)
            if self.device_outputs:
                out = {"main": dsbs, "depth": coded}
                if dmask is not None:
                    out["mask"] = dmask
                return out
            sbs.copy_(dsbs, non_blocking=True)
            if mask is not None:
                mask.copy_(dmask, non_blocking=True)
            hdepth = self._host_buf("depthcode", coded.shape)
            hdepth.copy_(coded, non_blocking=True)
            out = {"main": sbs, "depth": hdepth}
        if mask is not None:
            out["mask"] = deferred_mask if deferred_mask is not None else mask
        return out

    # ---- touchly1 -------------------------------------------------------------------------------------
    def _touchly1_chunk(self, depth_rgb, colour, first_frame):
        from .. import geometry as geo
        from .. import ops

        p, a = self.params, self.args
        n = depth_rgb.shape[0]
        d = self._dev_buf("d", depth_rgb.shape)
        c = self._dev_buf("c", colour.shape)
        d.copy_(depth_rgb, non_blocking=True)
        c.copy_(colour, non_blocking=True)
        out = self._dev_buf("t1", (n, 2 * self.h, self.w, 3))
        if p.transformations is None:  # fast path: no render pass (:548-552)
            out[:, :self.h].copy_(c)
            for k in range(n):
                scale = geo.master_fov_depth_scale(p.master_xfov, p.xfov_of(first_frame + k))
                ops.touchly_depth(d[k], a.touchly_min_depth, a.touchly_max_depth, False, p.max_depth, "D1", scale, out=out[k, self.h:])
        else:  # mono render at the frame's pose, then the rendered depth plane (:677-692)
            if self._zbuf is None:
                self._zbuf = ops.new_zbuf(1, self.w, self.h, self.device)
            rgb = self._dev_buf("mono", (n, self.h, self.w, 3))
            depth = self._dev_buf("monodepth", (n, self.h, self.w), torch.float32)
            sources, views = [], []
            for k in range(n):
                f = first_frame + k
                xf = p.xfov_of(f)
                K = geo.compute_camera_matrix(xf, None if p.xfovs is not None else p.yfov, self.w, self.h)
                sources.append(ops.make_source(self.w, self.h, K, p.max_depth, "D1", True, geo.master_fov_depth_scale(p.master_xfov, xf), False))
                views.append([ops.ViewSpec(np.asarray(p.transformations[f], dtype=np.float64), K[0, 0], K[1, 1], K[0, 2], K[1, 2])])
            ops.render_views(d, c, sources, views, self.w, self.h, self._zbuf, rgb, None, depth, p.bg_rgb, p.bg_rgb, 0, p.near)
            out[:, :self.h].copy_(rgb)
            for k in range(n):
                ops.touchly_depth(depth[k], a.touchly_min_depth, a.touchly_max_depth, True, decoder="F32", out=out[k, self.h:])
        if self.device_outputs:
            return {"main": out}
        host = self._host_buf("main", out.shape)
        host.copy_(out, non_blocking=True)
        return {"main": host}

    # ---- vr180 / touchly0 ---------------------------------------------------------------------------
    def _vr180_chunk(self, depth_rgb, colour, first_frame):
        """stereo_rerender.py:527-541,704-738,823-852,910-918: both eyes rendered by a square 1920x1920 camera of
        render_fov = max(75, widest source FOV), that FOV also taken as the master FOV of the depth scale; every
        panel (left, right, and for touchly0 the left eye's reverse-depth image) goes through the equirect remap."""
        import math

        from .. import geometry as geo
        from .. import ops, vr180

        p, a, side = self.params, self.args, self.vr_side
        n = depth_rgb.shape[0]
        d = self._dev_buf("d", depth_rgb.shape)
        c = self._dev_buf("c", colour.shape)
        d.copy_(depth_rgb, non_blocking=True)
        c.copy_(colour, non_blocking=True)
        if self._zbuf is None:
            self._zbuf = ops.new_zbuf(2, side, side, self.device)
        rect = self._dev_buf("rect", (n, side, 2 * side, 3))
        depth = self._dev_buf("rectdepth", (n, side, 2 * side), torch.float32) if self.touchly0 else None
        sources, views, fovs = [], [], []
        ipd = p.pupillary_distance / 1000
        for k in range(n):
            f = first_frame + k
            xf = p.xfov_of(f)
            K = geo.compute_camera_matrix(xf, None if p.xfovs is not None else p.yfov, self.w, self.h)
            fovx, fovy = geo.fov_from_camera_matrix(K)
            if max(fovx, fovy) >= 180:
                raise ValueError("fov cant be 180 or over, the tool is not built to handle fisheye distorted input video")
            render_fov = max(75, max(fovx, fovy))
            Kr = geo.compute_camera_matrix(render_fov, render_fov, side, side)
            scale = 1.0 / (math.tan(math.radians(render_fov / 2)) / math.tan(math.radians(xf / 2)))
            theta = None
            if p.convergence_depths is not None and float(p.convergence_depths[f]) != 0:
                theta = geo.convergence_angle(float(p.convergence_depths[f]) * scale, ipd)
            T = np.eye(4) if p.transformations is None else np.asarray(p.transformations[f], dtype=np.float64)
            sources.append(ops.make_source(self.w, self.h, K, p.max_depth, "D1", True, scale, False))
            views.append([ops.ViewSpec(geo.stereo_eye_pose(eye, ipd, theta) @ T, Kr[0, 0], Kr[1, 1], Kr[0, 2], Kr[1, 2]) for eye in ("left", "right")])
            fovs.append(render_fov)
        ops.render_views(d, c, sources, views, side, side, self._zbuf, rect, None, depth, p.bg_rgb, (0, 0, 0), 0, p.near)
        panels = 3 if self.touchly0 else 2
        out = self._dev_buf("vr", (n, side, panels * side, 3))
        tdepth = self._dev_buf("tdepth", (side, side, 3)) if self.touchly0 else None
        for k in range(n):
            mx, my = vr180.device_maps(side, side, fovs[k], self.device)
            for e in range(2):
                ops.remap_bilinear(rect[k, :, e * side:(e + 1) * side], mx, my, out=out[k, :, e * side:(e + 1) * side])
            if self.touchly0:
                ops.touchly_depth(depth[k, :, :side].contiguous(), a.touchly_min_depth, a.touchly_max_depth, True, decoder="F32", out=tdepth)
                ops.remap_bilinear(tdepth, mx, my, out=out[k, :, 2 * side:])
        if self.device_outputs:
            return {"main": out}
        host = self._host_buf("main", out.shape)
        host.copy_(out, non_blocking=True)
        return {"main": host}

    def render_chunk(self, depth_rgb: torch.Tensor, colour: torch.Tensor, first_frame: int) -> Dict[str, torch.Tensor]:
        if self.touchly1:
            return self._touchly1_chunk(depth_rgb, colour, first_frame)
        if self.vr180:
            return self._vr180_chunk(depth_rgb, colour, first_frame)
        return self._stereo_chunk(depth_rgb, colour, first_frame)


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
