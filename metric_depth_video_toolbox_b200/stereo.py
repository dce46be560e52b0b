"""Stereo re-render of a clip: the body of stereo_rerender.py's frame loop (:471-941) for whole batches
of frames, on the GPU.

`StereoRerenderer` is the public host API: it turns the script-level parameters (xfov / per-frame xfov
list, max_depth, pupillary distance, master FOV, optional smoothed convergence list, optional per-frame
camera poses) into per-frame constant blocks, picks the kernel path per clip

  * row-local fused kernel (`ops.stereo_rows`)      -- no pose file, no convergence rotation
  * virtual-source-row kernel (`ops.stereo_conv_rows(kernel="vrows")`) -- a convergence rotation without a pose file (what
    movie_2_3D runs), where the poses and the frame size pass the kernel's host-side limits check (`StereoParams.conv_kernel`)
  * generic frame loop (`ops.render_views`: colour-keyed splat + streaming resolve, two frames in flight on two streams)
    -- anything else (pose files, extreme convergence angles, widths that are not multiples of 32)

and runs it either on device-resident tensors (`render_device`) or on host arrays through a chunked,
triple-buffered H2D -> kernel -> D2H pipeline (`render_host`).  Every path writes the same bytes.  There is no CPU implementation.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import geometry as geo
from . import ops


@dataclass
class StereoParams:
    """Mirrors the stereo_rerender.py CLI values that reach the per-pixel path (:279-312)."""
    width: int
    height: int
    xfov: Optional[float] = None          # --xfov
    yfov: Optional[float] = None          # --yfov
    xfovs: Optional[Sequence[float]] = None  # --xfov_file (one value per frame; yfov then unused, :515-517)
    max_depth: float = 100                # --max_depth
    pupillary_distance: float = 63        # --pupillary_distance [mm]
    master_xfov: float = 45.0             # --master_xfov
    convergence_depths: Optional[Sequence[float]] = None  # already fill_nan + curve_fit'ed (:343-349)
    transformations: Optional[Sequence] = None            # per-frame 4x4, already re-based (:369-373)
    infill_mask: bool = True              # --infill_mask: green background, hole mask written
    mask_rgb: bool = False                # mask as the reference's green/black u8x3 image instead of u8
    near: float = geo.NEAR_PLANE
    force_generic: bool = False           # test / comparison aid: always take the K1+K2+K3 path
    conv_kernel: object = "auto"          # convergence without a pose file: "auto" = the virtual-source-row kernel (ops.stereo_conv_rows
                                          # kernel="vrows", no global z-buffer) where the poses and the frame size allow it, else the
                                          # two-lane generic loop; "vrows" = insist on it; True / "rows" = round 1's target-row kernel;
                                          # False / "generic" = always the generic loop.  Every choice writes the same bytes.

    def __post_init__(self):
        if self.xfov is None and self.yfov is None and self.xfovs is None:
            raise ValueError("Error: Either --xfov_file, --xfov or --yfov must be provided.")  # stereo_rerender.py:319-320

    @property
    def bg_rgb(self):
        return (0, 255, 0) if self.infill_mask else (0, 0, 0)  # stereo_rerender.py:554-556

    def xfov_of(self, frame: int) -> float:
        xf = self.xfovs[frame] if self.xfovs is not None else self.xfov
        if xf is None:
            # the reference dereferences xf / 2 with xf = None here (stereo_rerender.py:537): --yfov alone fails
            raise TypeError("unsupported operand: --xfov or --xfov_file is required by the master-FOV scaling")
        return float(xf)

    def row_local(self) -> bool:
        """True when every frame is a pure +-ipd/2 shift: v' == v, one fused kernel does it all."""
        return self.transformations is None and self.convergence_depths is None and not self.force_generic

    def conv_mode(self) -> str:
        """Which fused convergence kernel may run: "off" (force_generic; neither a convergence list nor a pose file: the row-local
        kernel's case; a pose file with anything but conv_kernel "auto"), else "auto" | "vrows" | "rows" | "generic" (conv_kernel)."""
        if self.force_generic or (self.transformations is None and self.convergence_depths is None):
            return "off"
        ck = self.conv_kernel
        if self.transformations is not None and ck != "auto":
            return "off"   # a pose file: only "auto" tries the virtual-row kernel (poses that are y-rotations + x-shifts qualify)
        if ck is True:
            ck = "rows"
        if ck is False or ck is None:
            ck = "generic"
        if ck not in ("auto", "vrows", "rows", "generic"):
            raise ValueError("conv_kernel is 'auto', 'vrows', 'rows' / True or 'generic' / False")
        if ck == "rows" and self.width > 4096:
            ck = "generic"
        return ck

    def conv_local(self) -> bool:
        """True when the eye poses are `rotation about y + shift along x` (convergence without a pose file) and a fused
        convergence kernel is allowed: the row displacement is then depth independent."""
        return self.conv_mode() in ("auto", "vrows", "rows")


class StereoRerenderer:
    def __init__(self, params: StereoParams, device: Optional[torch.device] = None):
        self.p = params
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._consts_cache = {}
        self._zbufs = {}  # one 2-view z-buffer per CUDA stream (render_host runs two streams)

    # ---- per-frame constants ---------------------------------------------------------------------
    def frame_constants(self, start: int, count: int) -> np.ndarray:
        """(count, 4) float32 rows of mdvt_stereo_frame for frames start .. start+count-1."""
        p = self.p
        if p.xfovs is None:
            row = ops.stereo_frame_constants(p.xfov_of(0), p.width, p.max_depth, p.pupillary_distance, p.master_xfov, p.near)
            return np.repeat(row[None], 1, axis=0)
        return np.stack([ops.stereo_frame_constants(p.xfov_of(f), p.width, p.max_depth, p.pupillary_distance, p.master_xfov, p.near)
                         for f in range(start, start + count)])

    def _device_constants(self, start: int, count: int) -> torch.Tensor:
        key = (start, count) if self.p.xfovs is not None else None
        if key not in self._consts_cache:
            if len(self._consts_cache) > 64:
                self._consts_cache.clear()
            self._consts_cache[key] = torch.from_numpy(self.frame_constants(start, count)).to(self.device)
        return self._consts_cache[key]

    def views_of(self, frame: int) -> List[ops.ViewSpec]:
        """Both eye cameras of one frame for the generic path (stereo_rerender.py:525,537-541,615-619,704-725,831-836)."""
        p = self.p
        xf = p.xfov_of(frame)
        K = geo.compute_camera_matrix(xf, None if p.xfovs is not None else p.yfov, p.width, p.height)
        scale = geo.master_fov_depth_scale(p.master_xfov, xf)
        ipd = p.pupillary_distance / 1000
        theta = None
        if p.convergence_depths is not None:
            conv = float(p.convergence_depths[frame])
            if conv != 0:  # "Convergence distance is zero, skipping convergence" (:711-713)
                theta = geo.convergence_angle(conv * scale, ipd)
        T = np.eye(4) if p.transformations is None else np.asarray(p.transformations[frame], dtype=np.float64)
        return [ops.ViewSpec(geo.stereo_eye_pose(eye, ipd, theta) @ T, K[0, 0], K[1, 1], K[0, 2], K[1, 2]) for eye in ("left", "right")]

    def packed_cameras(self, start: int, count: int):
        """The cameras of frames start .. start+count-1 for the generic path as packed arrays (ops.pack_sources /
        ops.pack_views): the same numbers as `views_of` / `ops.make_source` per frame, built with a handful of vectorised
        NumPy operations instead of ~40 us of Python per frame (which was 3/4 of the generic path's time per frame)."""
        import math

        p = self.p
        w, h = p.width, p.height
        frames = range(start, start + count)
        if p.xfovs is not None:
            Ks = [geo.compute_camera_matrix(p.xfov_of(f), None, w, h) for f in frames]
            scales = np.array([geo.master_fov_depth_scale(p.master_xfov, p.xfov_of(f)) for f in frames])
            fx, fy = np.array([K[0, 0] for K in Ks]), np.array([K[1, 1] for K in Ks])
            cx, cy = Ks[0][0, 2], Ks[0][1, 2]
        else:
            K = geo.compute_camera_matrix(p.xfov_of(start), p.yfov, w, h)
            fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
            scales = np.full(count, geo.master_fov_depth_scale(p.master_xfov, p.xfov_of(start)))
        ipd = p.pupillary_distance / 1000
        theta = np.zeros(count)
        if p.convergence_depths is not None:
            for k, f in enumerate(frames):
                conv = float(p.convergence_depths[f])
                if conv != 0:  # "Convergence distance is zero, skipping convergence" (:711-713)
                    theta[k] = geo.convergence_angle(conv * float(scales[k]), ipd)
        T = None if p.transformations is None else np.asarray([p.transformations[f] for f in frames], dtype=np.float64)
        M = np.zeros((count, 2, 3, 4))
        for e, sign in enumerate((1.0, -1.0)):   # left: Ry(-theta), +ipd/2; right: Ry(+theta), -ipd/2 (geo.stereo_eye_pose)
            ang = -sign * theta
            c = np.array([math.cos(a) if a else 1.0 for a in ang])[:, None]
            s = np.array([math.sin(a) if a else 0.0 for a in ang])[:, None]
            tx = sign * ipd / 2
            if T is None:
                M[:, e, 0, 0], M[:, e, 0, 2], M[:, e, 0, 3] = c[:, 0], s[:, 0], tx
                M[:, e, 1, 1] = 1.0
                M[:, e, 2, 0], M[:, e, 2, 2] = 0.0 - s[:, 0], c[:, 0]   # 0 - s: +0 where there is no rotation, like np.eye
            else:   # E @ T, summed in matmul's order; the terms that multiply an exact 0 are left out
                M[:, e, 0] = c * T[:, 0] + s * T[:, 2] + tx * T[:, 3]
                M[:, e, 1] = T[:, 1]
                M[:, e, 2] = -s * T[:, 0] + c * T[:, 2]
        bc = (lambda v: np.asarray(v)[:, None]) if p.xfovs is not None else (lambda v: v)
        views = ops.pack_views(M, bc(fx), bc(fy), cx, cy)
        sources = ops.pack_sources(w, h, fx, fy, cx, cy, p.max_depth, "D1", True, scales if p.xfovs is not None else scales[:1])
        return sources, views

    # ---- device-resident ------------------------------------------------------------------------------
    def render_device(self, depth_rgb: torch.Tensor, colour: torch.Tensor, start_frame: int = 0,
                      out_sbs: Optional[torch.Tensor] = None, out_mask: Optional[torch.Tensor] = None,
                      out_depth: Optional[torch.Tensor] = None, mask_rgb: Optional[bool] = None):
        """depth_rgb / colour: (n, H, W, 3) u8 CUDA.  Returns (sbs (n, H, 2W, 3), mask (n, H, 2W[, 3]) or None).
        `out_depth` (n, H, 2W) float32 receives the rendered depth of both eyes (0 where nothing was drawn)."""
        p = self.p
        n, h, w, _ = depth_rgb.shape
        if (w, h) != (p.width, p.height):
            raise ValueError(f"frames are {w}x{h}, parameters say {p.width}x{p.height}")
        mask_rgb = p.mask_rgb if mask_rgb is None else mask_rgb  # u8x3 background-colour mask image, or plain u8 {0,255}
        flags = (ops.FLAG_BG_COLLIDE if p.infill_mask else 0) | (ops.FLAG_MASK_RGB if mask_rgb else 0)
        if p.row_local():
            return ops.stereo_rows(depth_rgb, colour, self._device_constants(start_frame, n), p.bg_rgb, (0, 0, 0), flags,
                                   out_sbs, out_mask, want_mask=p.infill_mask or out_mask is not None, out_depth=out_depth)
        if out_sbs is None:
            out_sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device=depth_rgb.device)
        if out_mask is None and p.infill_mask:
            out_mask = torch.empty((n, h, 2 * w) + ((3,) if mask_rgb else ()), dtype=torch.uint8, device=depth_rgb.device)
        mode = p.conv_mode()
        if mode in ("auto", "vrows", "rows"):  # convergence only: fused kernels without a global z-buffer
            key = ("conv", start_frame, n)
            if key not in self._consts_cache:
                if len(self._consts_cache) > 64:
                    self._consts_cache.clear()
                try:
                    host = ops.conv_frames_packed(*self.packed_cameras(start_frame, n), p.near)
                    self._consts_cache[key] = (torch.from_numpy(host).to(self.device), ops.conv_vrows_supported(host, w, h))
                except ValueError:   # a pose of the chunk is not `y rotation + x shift` (pose files): the generic loop
                    if p.transformations is None:
                        raise
                    self._consts_cache[key] = (None, False)
            frames_dev, vrows_ok = self._consts_cache[key]
            aligned = all(t is None or t.data_ptr() % 16 == 0 for t in (depth_rgb, colour, out_sbs, out_mask, out_depth))
            if mode == "vrows" and not (vrows_ok and aligned):
                raise ValueError("conv_kernel='vrows': the poses / frame size / buffer alignment are outside the virtual-row kernel's limits")
            if mode != "rows" and vrows_ok and aligned:
                return ops.stereo_conv_rows(depth_rgb, colour, frames_dev, p.bg_rgb, (0, 0, 0), flags, out_sbs, out_mask,
                                            want_mask=False, out_depth=out_depth, kernel="vrows")
            if mode == "rows":
                return ops.stereo_conv_rows(depth_rgb, colour, frames_dev, p.bg_rgb, (0, 0, 0), flags, out_sbs, out_mask,
                                            want_mask=False, out_depth=out_depth)
        sources, views = self.packed_cameras(start_frame, n)
        # generic path: per frame K1+K2 into a persistent 2-view z-buffer, K3 for both eyes straight into the SBS halves
        zkey = torch.cuda.current_stream(depth_rgb.device).cuda_stream
        zbuf = self._zbufs.get(zkey)
        sets = 1 if os.environ.get("MDVT_ZBUF_SETS", "2") == "1" else 2   # two sets of planes: frames alternate between two streams
        if zbuf is None or tuple(zbuf.shape) != (2 * sets, h, w) or zbuf.device != depth_rgb.device:
            zbuf = self._zbufs[zkey] = ops.new_zbuf(2 * sets, w, h, depth_rgb.device)
        ops.render_views(depth_rgb, colour, sources, views, w, h, zbuf, out_sbs, out_mask, out_depth, p.bg_rgb, (0, 0, 0), flags, p.near)
        return out_sbs, out_mask

    # ---- host arrays, pipelined -----------------------------------------------------------------------
    def render_host(self, depth_rgb, colour, out_sbs=None, out_mask=None, start_frame: int = 0, chunk_frames: int = 4,
                    mask_format: str = "u8"):
        """depth_rgb / colour: (n, H, W, 3) u8 host arrays (NumPy or CPU tensors; pinned memory makes the
        copies asynchronous).  Frames stream through three `chunk_frames`-sized device staging buffers, each with a CUDA stream
        of its own, so H2D, the kernels and D2H overlap (4 frames per buffer measured best on the B200: 3.3-3.4 k frames/s at
        1080p against 3.07 k with 8 and 2.7 k with 16; a separate upload / kernel / download stream design measured the same:
        profiles/r02_host_pipeline_sweep*.txt).  Returns host tensors (sbs, mask); it waits for the last copy,
        so the returned buffers are complete.
        mask_format "bits": the hole mask crosses PCIe as one bit per pixel -- (n, H, 2W/8) u8, most significant bit
        first; `ops.unpack_mask_bits` (numpy.unpackbits) gives the u8 {0, 255} plane back.  It takes the mask's share of the
        device-to-host bytes from 25 % to 4 % (the host link, not the kernel, bounds this call); plain u8 masks only."""
        p = self.p
        d_host = torch.as_tensor(depth_rgb)
        c_host = torch.as_tensor(colour)
        n, h, w, _ = d_host.shape
        bits = mask_format == "bits"
        if mask_format not in ("u8", "bits") or (bits and (p.mask_rgb or not p.infill_mask or (2 * w) % 8)):
            raise ValueError("mask_format is 'u8' or 'bits' (bits: u8 masks of a width that is a multiple of 8)")
        if out_sbs is None:
            out_sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, pin_memory=True)
        if out_mask is None and p.infill_mask:
            out_mask = torch.empty((n, h, 2 * w // 8) if bits else (n, h, 2 * w) + ((3,) if p.mask_rgb else ()), dtype=torch.uint8, pin_memory=True)
        out_sbs_t, out_mask_t = torch.as_tensor(out_sbs), (None if out_mask is None else torch.as_tensor(out_mask))
        chunk = max(1, min(int(os.environ.get("MDVT_HOST_CHUNK", chunk_frames)), n))   # tuning aids: frames per staging buffer,
        n_slots = max(2, int(os.environ.get("MDVT_HOST_SLOTS", "3")))                     # staging buffers / streams in flight
        slots = self._host_pipeline_slots(n_slots, chunk, h, w)
        caller = torch.cuda.current_stream(self.device)
        for s in slots:
            s["stream"].wait_stream(caller)
        for i, f0 in enumerate(range(0, n, chunk)):
            s = slots[i % n_slots]
            cnt = min(chunk, n - f0)
            with torch.cuda.stream(s["stream"]):
                s["d"][:cnt].copy_(d_host[f0:f0 + cnt], non_blocking=True)
                s["c"][:cnt].copy_(c_host[f0:f0 + cnt], non_blocking=True)
                self.render_device(s["d"][:cnt], s["c"][:cnt], start_frame + f0, s["sbs"][:cnt],
                                   None if s["mask"] is None else s["mask"][:cnt])
                out_sbs_t[f0:f0 + cnt].copy_(s["sbs"][:cnt], non_blocking=True)
                if out_mask_t is not None and bits:
                    if s.get("bits") is None:
                        s["bits"] = torch.empty((chunk, h, 2 * w // 8), dtype=torch.uint8, device=self.device)
                    out_mask_t[f0:f0 + cnt].copy_(ops.pack_mask_bits(s["mask"][:cnt], s["bits"][:cnt]), non_blocking=True)
                elif out_mask_t is not None:
                    out_mask_t[f0:f0 + cnt].copy_(s["mask"][:cnt], non_blocking=True)
        for s in slots:
            caller.wait_stream(s["stream"])
            s["stream"].synchronize()   # "returns host tensors": the asynchronous D2H copies have landed when this returns
        return out_sbs, out_mask

    def _host_pipeline_slots(self, n_slots, chunk, h, w):
        key = (n_slots, chunk, h, w, self.p.mask_rgb, self.p.infill_mask)
        if getattr(self, "_slots_key", None) != key:
            dev = self.device
            mask_shape = (chunk, h, 2 * w) + ((3,) if self.p.mask_rgb else ())
            self._slots = [dict(stream=torch.cuda.Stream(device=dev),
                                d=torch.empty((chunk, h, w, 3), dtype=torch.uint8, device=dev),
                                c=torch.empty((chunk, h, w, 3), dtype=torch.uint8, device=dev),
                                sbs=torch.empty((chunk, h, 2 * w, 3), dtype=torch.uint8, device=dev),
                                mask=torch.empty(mask_shape, dtype=torch.uint8, device=dev) if self.p.infill_mask else None)
                           for _ in range(n_slots)]
            self._slots_key = key
        return self._slots
