"""Tensor-in / tensor-out wrappers over the C ABI (include/mdvt_b200.h).

PyTorch is used for device memory and streams only: every function takes CUDA tensors, enqueues
one hand-written kernel on the current stream through ctypes and returns CUDA tensors.  There is
no CPU implementation behind these calls.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import DECODERS, FLAG_ANYWIDTH, FLAG_BG_COLLIDE, FLAG_MASK_RGB, FLAG_RESET_ZBUF, MdvtError  # noqa: F401

FULL_SCALE = 255 ** 4
NEAR_PLANE = 1e-4  # depth_map_tools.py:1520


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def pack_rgb(rgb: Sequence[int]) -> int:
    r, g, b = (int(v) & 0xFF for v in rgb)
    return r | (g << 8) | (b << 16)


def dec_const(max_depth, decoder: str = "D1") -> float:
    """The one float32 constant of the decode."""
    if decoder == "D1":
        return float(np.float32(float(max_depth) / FULL_SCALE))
    return float(np.float32(FULL_SCALE / max_depth))


# ---------------------------------------------------------------------------------------------
# wire-format codec
# ---------------------------------------------------------------------------------------------
def decode_depth(rgb: torch.Tensor, max_depth, bit16: bool = True, decoder: str = "D1", want_codes: bool = False,
                 want_depth: bool = True):
    """(..., 3) u8 RGB-order -> float32 depth (and/or uint32 codes stored in an int32 tensor's bits)."""
    _need(rgb, torch.uint8, "rgb")
    if rgb.shape[-1] != 3:
        raise ValueError("rgb must have a trailing dimension of 3")
    shape = rgb.shape[:-1]
    n = rgb.numel() // 3
    depth = torch.empty(shape, dtype=torch.float32, device=rgb.device) if want_depth else None
    codes = torch.empty(shape, dtype=torch.uint32, device=rgb.device) if want_codes else None
    lib = _lib.load()
    _lib.check(lib.mdvt_decode_depth(_ptr(rgb), n, DECODERS[decoder], int(bool(bit16)), dec_const(max_depth, decoder),
                                     _ptr(codes), _ptr(depth), _stream()))
    if want_codes and want_depth:
        return depth, codes
    return codes if want_codes else depth


def encode_depth(depth: torch.Tensor, max_depth, bit16: bool = True, bgr_order: bool = True, want_codes: bool = False):
    """float32 metres -> (..., 3) u8 (B,G,R by default, as the reference hands to cv2)."""
    _need(depth, torch.float32, "depth")
    n = depth.numel()
    pix = torch.empty(depth.shape + (3,), dtype=torch.uint8, device=depth.device)
    codes = torch.empty(depth.shape, dtype=torch.uint32, device=depth.device) if want_codes else None
    lib = _lib.load()
    _lib.check(lib.mdvt_encode_depth(_ptr(depth), n, float(max_depth), int(bool(bit16)), int(bool(bgr_order)), _ptr(codes),
                                     _ptr(pix), _stream()))
    return (pix, codes) if want_codes else pix


# ---------------------------------------------------------------------------------------------
# geometry
# ---------------------------------------------------------------------------------------------
def make_source(width: int, height: int, K: np.ndarray, max_depth, decoder: str = "D1", bit16: bool = True,
                depth_scale: float = 1.0, of_by_one: bool = False) -> _lib.Source:
    s = _lib.Source()
    s.width, s.height, s.decoder, s.bit16 = int(width), int(height), DECODERS[decoder], int(bool(bit16))
    s.dec_const = dec_const(max_depth, decoder)
    s.depth_scale = float(np.float32(depth_scale))
    s.fx, s.fy, s.cx, s.cy = (float(np.float32(v)) for v in (K[0][0], K[1][1], K[0][2], K[1][2]))
    s.grid_sx = float(np.float32((width + 1) / width)) if of_by_one else 1.0
    s.grid_sy = float(np.float32((height + 1) / height)) if of_by_one else 1.0
    return s


def _pose12(pose, ctype):
    if pose is None:
        return None
    m = np.asarray(pose, dtype=np.float64)[:3, :4].reshape(-1)
    return (ctype * 12)(*[float(v) for v in m])


def unproject(depth_rgb: torch.Tensor, source: _lib.Source, pose=None, dtype=torch.float32, K: Optional[np.ndarray] = None):
    """(H, W, 3) u8 -> (H*W, 3) points.  float64 needs the float64 intrinsics `K` (3x3)."""
    _need(depth_rgb, torch.uint8, "depth_rgb")
    if tuple(depth_rgb.shape) != (source.height, source.width, 3):
        raise ValueError(f"depth_rgb shape {tuple(depth_rgb.shape)} != ({source.height}, {source.width}, 3)")
    n = source.width * source.height
    out = torch.empty((n, 3), dtype=dtype, device=depth_rgb.device)
    lib = _lib.load()
    if dtype == torch.float32:
        _lib.check(lib.mdvt_unproject_f32(_ptr(depth_rgb), C.byref(source), _pose12(pose, C.c_float), _ptr(out), _stream()))
    elif dtype == torch.float64:
        if K is None:
            raise ValueError("float64 unprojection needs the float64 camera matrix K")
        k4 = (C.c_double * 4)(float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]))
        _lib.check(lib.mdvt_unproject_f64(_ptr(depth_rgb), C.byref(source), k4, _pose12(pose, C.c_double), _ptr(out), _stream()))
    else:
        raise TypeError("dtype must be torch.float32 or torch.float64")
    return out


@dataclass
class ViewSpec:
    """One virtual camera: 3x4 (or 4x4) pose source-camera -> view-camera, output intrinsics."""
    M: np.ndarray
    fx: float
    fy: float
    cx: float
    cy: float

    def to_c(self) -> _lib.View:
        v = _lib.View()
        m = np.asarray(self.M, dtype=np.float64)[:3, :4].astype(np.float32).reshape(-1)
        for k in range(12):
            v.M[k] = float(m[k])
        v.fx, v.fy, v.cx, v.cy = (float(np.float32(x)) for x in (self.fx, self.fy, self.cx, self.cy))
        return v


def new_zbuf(n_views: int, out_w: int, out_h: int, device) -> torch.Tensor:
    z = torch.empty((n_views, out_h, out_w), dtype=torch.int64, device=device)
    zbuf_clear(z)
    return z


def zbuf_clear(zbuf: torch.Tensor):
    _need(zbuf, torch.int64, "zbuf")
    _lib.check(_lib.load().mdvt_zbuf_clear(_ptr(zbuf), zbuf.numel(), _stream()))


def project_splat(depth_rgb: torch.Tensor, source: _lib.Source, views: Sequence[ViewSpec], out_w: int, out_h: int,
                  zbuf: torch.Tensor, near: float = NEAR_PLANE, want_uvz: bool = False):
    """K1+K2: merge every source pixel into `zbuf` (n_views, out_h, out_w) int64 (u64 bits)."""
    _need(depth_rgb, torch.uint8, "depth_rgb")
    _need(zbuf, torch.int64, "zbuf")
    if tuple(depth_rgb.shape) != (source.height, source.width, 3):
        raise ValueError("depth_rgb shape does not match the source description")
    if tuple(zbuf.shape) != (len(views), out_h, out_w):
        raise ValueError(f"zbuf shape {tuple(zbuf.shape)} != ({len(views)}, {out_h}, {out_w})")
    n = source.width * source.height
    uvz = torch.empty((len(views), n, 3), dtype=torch.float32, device=depth_rgb.device) if want_uvz else None
    arr = (_lib.View * len(views))(*[v.to_c() for v in views])
    _lib.check(_lib.load().mdvt_project_splat(_ptr(depth_rgb), C.byref(source), arr, len(views), float(np.float32(near)),
                                              int(out_w), int(out_h), _ptr(zbuf), _ptr(uvz), _stream()))
    return uvz


def resolve(zbuf_view: torch.Tensor, colour: torch.Tensor, bg_rgb=(0, 0, 0), fill_rgb=(0, 0, 0), flags: int = 0,
            out_rgb: Optional[torch.Tensor] = None, out_mask: Optional[torch.Tensor] = None, want_depth: bool = False,
            want_ids: bool = False):
    """K3 for one view.  `out_rgb` (H, W, 3) / `out_mask` (H, W[, 3]) may be column slices of wider
    side-by-side tensors (row stride is passed through as the pitch)."""
    _need(zbuf_view, torch.int64, "zbuf_view")
    _need(colour, torch.uint8, "colour")
    out_h, out_w = zbuf_view.shape
    dev = zbuf_view.device
    mask_bpp = 3 if flags & FLAG_MASK_RGB else 1
    if out_rgb is None:
        out_rgb = torch.empty((out_h, out_w, 3), dtype=torch.uint8, device=dev)
    if out_mask is None:
        out_mask = torch.empty((out_h, out_w) + ((3,) if mask_bpp == 3 else ()), dtype=torch.uint8, device=dev)
    for t, bpp, name in ((out_rgb, 3, "out_rgb"), (out_mask, mask_bpp, "out_mask")):
        if t.dtype != torch.uint8 or not t.is_cuda or t.shape[0] != out_h or t.shape[1] != out_w:
            raise ValueError(f"{name} has the wrong dtype / device / shape")
        inner_ok = (t.dim() == 2 and bpp == 1 and t.stride(1) == 1) or (t.dim() == 3 and t.shape[2] == bpp and t.stride(2) == 1 and t.stride(1) == bpp)
        if not inner_ok:
            raise ValueError(f"{name} rows must be dense")
    depth = torch.empty((out_h, out_w), dtype=torch.float32, device=dev) if want_depth else None
    ids = torch.empty((out_h, out_w), dtype=torch.int32, device=dev) if want_ids else None
    _lib.check(_lib.load().mdvt_resolve(_ptr(zbuf_view), _ptr(colour), out_w, out_h, pack_rgb(bg_rgb), pack_rgb(fill_rgb), flags,
                                        _ptr(out_rgb), out_rgb.stride(0), _ptr(out_mask), out_mask.stride(0), _ptr(depth), _ptr(ids),
                                        _stream()))
    return out_rgb, out_mask, depth, ids


# ---------------------------------------------------------------------------------------------
# row-local stereo fast path
# ---------------------------------------------------------------------------------------------
def stereo_frame_constants(xfov_deg: float, width: int, max_depth, pupillary_distance_mm: float, master_xfov_deg: float,
                           near: float = NEAR_PLANE) -> np.ndarray:
    """The four float32 constants of mdvt_stereo_frame for one frame.  fx: depth_map_tools.py:902-934;
    master-FOV scale: stereo_rerender.py:537-538; ipd: :458-459."""
    fx = width / (2 * np.tan(np.deg2rad(xfov_deg) / 2))
    scale = 1.0 / (math.tan(math.radians(master_xfov_deg / 2)) / math.tan(math.radians(xfov_deg / 2)))
    ipd = pupillary_distance_mm / 1000
    return np.array([dec_const(max_depth, "D1"), np.float32(scale), np.float32(fx * ipd / 2), np.float32(near)], dtype=np.float32)


def stereo_rows(depth_rgb: torch.Tensor, colour: torch.Tensor, frames: torch.Tensor, bg_rgb=(0, 0, 0), fill_rgb=(0, 0, 0),
                flags: int = 0, out_sbs: Optional[torch.Tensor] = None, out_mask: Optional[torch.Tensor] = None,
                want_mask: bool = True):
    """Fused stereo kernel over a batch: depth_rgb / colour (n, H, W, 3) u8; frames (n, 4) or (1, 4) float32
    CUDA tensor of mdvt_stereo_frame rows.  Returns (sbs (n, H, 2W, 3), mask (n, H, 2W[, 3]) or None)."""
    _need(depth_rgb, torch.uint8, "depth_rgb")
    _need(colour, torch.uint8, "colour")
    _need(frames, torch.float32, "frames")
    if depth_rgb.dim() != 4 or depth_rgb.shape[-1] != 3 or depth_rgb.shape != colour.shape:
        raise ValueError("depth_rgb and colour must both be (n, H, W, 3)")
    n, h, w, _ = depth_rgb.shape
    if frames.dim() != 2 or frames.shape[1] != 4 or frames.shape[0] not in (1, n):
        raise ValueError("frames must be (n, 4) or (1, 4) float32")
    dev = depth_rgb.device
    mask_bpp = 3 if flags & FLAG_MASK_RGB else 1
    if out_sbs is None:
        out_sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device=dev)
    _need(out_sbs, torch.uint8, "out_sbs")
    if out_sbs.numel() != n * h * 2 * w * 3:
        raise ValueError("out_sbs has the wrong size")
    if want_mask and out_mask is None:
        out_mask = torch.empty((n, h, 2 * w) + ((3,) if mask_bpp == 3 else ()), dtype=torch.uint8, device=dev)
    if out_mask is not None:
        _need(out_mask, torch.uint8, "out_mask")
        if out_mask.numel() != n * h * 2 * w * mask_bpp:
            raise ValueError("out_mask has the wrong size")
    _lib.check(_lib.load().mdvt_stereo_rows(_ptr(depth_rgb), _ptr(colour), n, w, h, _ptr(frames), int(frames.shape[0] == n and n > 1),
                                            pack_rgb(bg_rgb), pack_rgb(fill_rgb), flags, _ptr(out_sbs), _ptr(out_mask), _stream()))
    return out_sbs, out_mask
