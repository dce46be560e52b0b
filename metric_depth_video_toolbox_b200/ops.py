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
    """float32 (or float64: clipped and scaled in float64 like the reference does for such an array) metres -> (..., 3) u8
    (B,G,R by default, as the reference hands to cv2)."""
    if depth.dtype != torch.float64:
        _need(depth, torch.float32, "depth")
    else:
        _need(depth, torch.float64, "depth")
    n = depth.numel()
    pix = torch.empty(depth.shape + (3,), dtype=torch.uint8, device=depth.device)
    codes = torch.empty(depth.shape, dtype=torch.uint32, device=depth.device) if want_codes else None
    lib = _lib.load()
    fn = lib.mdvt_encode_depth_f64 if depth.dtype == torch.float64 else lib.mdvt_encode_depth
    _lib.check(fn(_ptr(depth), n, float(max_depth), int(bool(bit16)), int(bool(bgr_order)), _ptr(codes), _ptr(pix), _stream()))
    return (pix, codes) if want_codes else pix


def codes_to_depth(codes: torch.Tensor, max_depth, decoder: str = "D1") -> torch.Tensor:
    """decode_uint32_as_depth (depth_frames_helper.py:13-24): uint32 codes -> float32 metres."""
    _need(codes, torch.uint32, "codes")
    out = torch.empty(codes.shape, dtype=torch.float32, device=codes.device)
    _lib.check(_lib.load().mdvt_codes_to_depth(_ptr(codes), codes.numel(), DECODERS[decoder], dec_const(max_depth, decoder), _ptr(out),
                                               _stream()))
    return out


def codes_to_pixels(codes: torch.Tensor, bit16: bool = False, bgr_order: bool = True) -> torch.Tensor:
    """encode_data_as_BGR (depth_frames_helper.py:48-61): uint32 plane -> (..., 3) u8."""
    _need(codes, torch.uint32, "codes")
    out = torch.empty(codes.shape + (3,), dtype=torch.uint8, device=codes.device)
    _lib.check(_lib.load().mdvt_codes_to_pixels(_ptr(codes), codes.numel(), int(bool(bit16)), int(bool(bgr_order)), _ptr(out), _stream()))
    return out


def _depth_source(depth_src: torch.Tensor, decoder: str):
    """(tensor, pixel count, shape without the channel axis) of a wire-format (..., 3) u8 or float32 depth tensor."""
    if decoder == "F32":
        _need(depth_src, torch.float32, "depth_src")
        return depth_src.numel(), tuple(depth_src.shape)
    _need(depth_src, torch.uint8, "depth_src")
    if depth_src.shape[-1] != 3:
        raise ValueError("a wire-format frame must have a trailing dimension of 3")
    return depth_src.numel() // 3, tuple(depth_src.shape[:-1])


def depth_to_grey(depth_src: torch.Tensor, max_depth, bits: int, decoder: str = "D2", bit16: bool = True) -> torch.Tensor:
    """convert_metric_depth_video_to_other_format.py:752-760: --bit16 -> (...,) uint16, --bit8 -> (..., 3) u8."""
    n, shape = _depth_source(depth_src, decoder)
    if bits == 16:
        factor, out = (255 ** 2) / max_depth, torch.empty(shape, dtype=torch.uint16, device=depth_src.device)
    elif bits == 8:
        factor, out = 255 / max_depth, torch.empty(shape + (3,), dtype=torch.uint8, device=depth_src.device)
    else:
        raise ValueError("bits must be 8 or 16")
    dc = 1.0 if decoder == "F32" else dec_const(max_depth, decoder)
    _lib.check(_lib.load().mdvt_depth_to_grey(_ptr(depth_src), n, DECODERS[decoder], int(bool(bit16)), dc, float(np.float32(factor)), bits,
                                              3 if bits == 8 else 1, _ptr(out), _stream()))
    return out


def touchly_depth(depth_src: torch.Tensor, touchly_min: float, touchly_max: float, zero_is_far: bool, max_depth=100,
                  decoder: str = "D1", depth_scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One frame (H, W, 3) u8 wire format or (H, W) f32 -> Touchly reverse-depth image (H, W, 3) u8
    (stereo_rerender.py:548-552,687-690,826-829).  `out` may be a row/column slice of a larger frame."""
    _, shape = _depth_source(depth_src, decoder)
    if len(shape) != 2:
        raise ValueError("touchly_depth takes one frame")
    h, w = shape
    if out is None:
        out = torch.empty((h, w, 3), dtype=torch.uint8, device=depth_src.device)
    if out.dtype != torch.uint8 or not out.is_cuda or tuple(out.shape) != (h, w, 3) or out.stride(2) != 1 or out.stride(1) != 3:
        raise ValueError("out must be a (H, W, 3) u8 CUDA view with dense rows")
    dc = 1.0 if decoder == "F32" else dec_const(max_depth, decoder)
    _lib.check(_lib.load().mdvt_touchly_depth(_ptr(depth_src), w, h, DECODERS[decoder], 1, dc, float(np.float32(depth_scale)),
                                              float(np.float32(touchly_min)), float(np.float32(touchly_max)),
                                              float(np.float32(255 / (touchly_max - touchly_min))), int(bool(zero_is_far)),
                                              _ptr(out), out.stride(0), _stream()))
    return out


def remap_bilinear(src: torch.Tensor, map_x: torch.Tensor, map_y: torch.Tensor, out: Optional[torch.Tensor] = None,
                   border_rgb=(0, 0, 0)) -> torch.Tensor:
    """cv2.remap(src, map_x, map_y, INTER_LINEAR, BORDER_CONSTANT) for (H, W, 3) u8 images, bit-exact with OpenCV.
    `src` / `out` may be column slices of wider frames (dense rows)."""
    for t, name in ((src, "src"), (out, "out")):
        if t is None:
            continue
        if t.dtype != torch.uint8 or not t.is_cuda or t.dim() != 3 or t.shape[2] != 3 or t.stride(2) != 1 or t.stride(1) != 3:
            raise ValueError(f"{name} must be a (H, W, 3) u8 CUDA view with dense rows")
    _need(map_x, torch.float32, "map_x")
    _need(map_y, torch.float32, "map_y")
    if map_x.shape != map_y.shape or map_x.dim() != 2:
        raise ValueError("map_x / map_y must be (H_out, W_out) float32")
    dh, dw = map_x.shape
    if out is None:
        out = torch.empty((dh, dw, 3), dtype=torch.uint8, device=src.device)
    if tuple(out.shape) != (dh, dw, 3):
        raise ValueError("out does not match the maps")
    _lib.check(_lib.load().mdvt_remap_bilinear_u8x3(_ptr(src), src.shape[1], src.shape[0], src.stride(0), _ptr(map_x), _ptr(map_y), dw, dh,
                                                    pack_rgb(border_rgb), _ptr(out), out.stride(0), _stream()))
    return out


# ---------------------------------------------------------------------------------------------
# geometry
# ---------------------------------------------------------------------------------------------
def make_source(width: int, height: int, K: np.ndarray, max_depth=100, decoder: str = "D1", bit16: bool = True,
                depth_scale: float = 1.0, of_by_one: bool = False) -> _lib.Source:
    """decoder "F32": the source tensor is a (H, W) float32 depth plane instead of a wire-format frame."""
    s = _lib.Source()
    s.width, s.height, s.decoder, s.bit16 = int(width), int(height), DECODERS[decoder], int(bool(bit16))
    s.dec_const = 1.0 if decoder == "F32" else dec_const(max_depth, decoder)
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


def _need_source(depth_src: torch.Tensor, source: _lib.Source) -> torch.Tensor:
    """The frame a mdvt_source describes: (H, W, 3) u8 wire format, or (H, W) float32 for decoder F32."""
    if source.decoder == _lib.SOURCE_F32:
        _need(depth_src, torch.float32, "depth_src")
        want = (source.height, source.width)
    else:
        _need(depth_src, torch.uint8, "depth_src")
        want = (source.height, source.width, 3)
    if tuple(depth_src.shape) != want:
        raise ValueError(f"depth_src shape {tuple(depth_src.shape)} != {want}")
    return depth_src


def _k4(K):
    if K is None:
        raise ValueError("the float64 path needs the float64 camera matrix K")
    return (C.c_double * 4)(float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]))


def unproject(depth_rgb: torch.Tensor, source: _lib.Source, pose=None, dtype=torch.float32, K: Optional[np.ndarray] = None):
    """(H, W, 3) u8 (or (H, W) f32 for an F32 source) -> (H*W, 3) points.  float64 needs the float64 intrinsics `K` (3x3)."""
    _need_source(depth_rgb, source)
    n = source.width * source.height
    out = torch.empty((n, 3), dtype=dtype, device=depth_rgb.device)
    lib = _lib.load()
    if dtype == torch.float32:
        _lib.check(lib.mdvt_unproject_f32(_ptr(depth_rgb), C.byref(source), _pose12(pose, C.c_float), _ptr(out), _stream()))
    elif dtype == torch.float64:
        k4 = _k4(K)
        _lib.check(lib.mdvt_unproject_f64(_ptr(depth_rgb), C.byref(source), k4, _pose12(pose, C.c_double), _ptr(out), _stream()))
    else:
        raise TypeError("dtype must be torch.float32 or torch.float64")
    return out


def transform_points(xyz: torch.Tensor, transform, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """depth_map_tools.transform_points (:977-1004) on (N, 3) float64 device points."""
    _need(xyz, torch.float64, "xyz")
    if xyz.dim() != 2 or xyz.shape[1] != 3:
        raise ValueError("xyz must be (N, 3)")
    out = torch.empty_like(xyz) if out is None else _need(out, torch.float64, "out")
    _lib.check(_lib.load().mdvt_transform_points_f64(_ptr(xyz), xyz.shape[0], _pose12(transform, C.c_double), _ptr(out), _stream()))
    return out


def project_points(xyz: torch.Tensor, K) -> torch.Tensor:
    """depth_map_tools.project_3d_points_to_2d (:1057-1060) on (N, 3) float64 device points -> (N, 2)."""
    _need(xyz, torch.float64, "xyz")
    if xyz.dim() != 2 or xyz.shape[1] != 3:
        raise ValueError("xyz must be (N, 3)")
    out = torch.empty((xyz.shape[0], 2), dtype=torch.float64, device=xyz.device)
    _lib.check(_lib.load().mdvt_project_points_f64(_ptr(xyz), xyz.shape[0], _k4(K), _ptr(out), _stream()))
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


SOURCE_DTYPE = np.dtype([("width", "<i4"), ("height", "<i4"), ("decoder", "<i4"), ("bit16", "<i4"), ("dec_const", "<f4"),
                         ("depth_scale", "<f4"), ("fx", "<f4"), ("fy", "<f4"), ("cx", "<f4"), ("cy", "<f4"), ("grid_sx", "<f4"),
                         ("grid_sy", "<f4")])   # mdvt_source, field for field


def pack_views(M, fx, fy, cx, cy) -> np.ndarray:
    """Cameras as rows of mdvt_view without a Python object per camera: M (..., 3|4, 4) float64 poses, intrinsics
    broadcastable to M's leading shape -> (..., 16) float32 [M row-major 3x4, fx, fy, cx, cy] (what ViewSpec.to_c writes)."""
    M = np.asarray(M, dtype=np.float64)[..., :3, :4]
    lead = M.shape[:-2]
    out = np.empty(lead + (16,), dtype=np.float32)
    out[..., :12] = M.reshape(lead + (12,)).astype(np.float32)
    for k, v in enumerate((fx, fy, cx, cy)):
        out[..., 12 + k] = np.asarray(v, dtype=np.float64).astype(np.float32)
    return out


def pack_sources(width: int, height: int, fx, fy, cx, cy, max_depth=100, decoder: str = "D1", bit16: bool = True, depth_scale=1.0,
                 of_by_one: bool = False) -> np.ndarray:
    """n mdvt_source records at once (make_source for per-frame intrinsics / depth scales): fx, fy, cx, cy, depth_scale are
    scalars or (n,) arrays -> structured array of SOURCE_DTYPE."""
    n = int(np.broadcast(fx, fy, cx, cy, depth_scale).size)
    out = np.zeros(n, dtype=SOURCE_DTYPE)
    out["width"], out["height"], out["decoder"], out["bit16"] = int(width), int(height), DECODERS[decoder], int(bool(bit16))
    out["dec_const"] = 1.0 if decoder == "F32" else dec_const(max_depth, decoder)
    out["depth_scale"] = np.asarray(depth_scale, dtype=np.float64).astype(np.float32)
    for k, v in (("fx", fx), ("fy", fy), ("cx", cx), ("cy", cy)):
        out[k] = np.asarray(v, dtype=np.float64).astype(np.float32)
    out["grid_sx"] = np.float32((width + 1) / width) if of_by_one else 1.0
    out["grid_sy"] = np.float32((height + 1) / height) if of_by_one else 1.0
    return out


def new_zbuf(n_views: int, out_w: int, out_h: int, device) -> torch.Tensor:
    z = torch.empty((n_views, out_h, out_w), dtype=torch.int64, device=device)
    zbuf_clear(z)
    return z


def zbuf_clear(zbuf: torch.Tensor):
    _need(zbuf, torch.int64, "zbuf")
    _lib.check(_lib.load().mdvt_zbuf_clear(_ptr(zbuf), zbuf.numel(), _stream()))


def project_splat(depth_rgb: torch.Tensor, source: _lib.Source, views: Sequence[ViewSpec], out_w: int, out_h: int,
                  zbuf: torch.Tensor, near: float = NEAR_PLANE, want_uvz: bool = False, id_offset: int = 0):
    """K1+K2: merge every source pixel into `zbuf` (n_views, out_h, out_w) int64 (u64 bits)."""
    _need_source(depth_rgb, source)
    _need(zbuf, torch.int64, "zbuf")
    if tuple(zbuf.shape) != (len(views), out_h, out_w):
        raise ValueError(f"zbuf shape {tuple(zbuf.shape)} != ({len(views)}, {out_h}, {out_w})")
    n = source.width * source.height
    uvz = torch.empty((len(views), n, 3), dtype=torch.float32, device=depth_rgb.device) if want_uvz else None
    arr = (_lib.View * len(views))(*[v.to_c() for v in views])
    _lib.check(_lib.load().mdvt_project_splat(_ptr(depth_rgb), C.byref(source), arr, len(views), float(np.float32(near)),
                                              int(out_w), int(out_h), int(id_offset), _ptr(zbuf), _ptr(uvz), _stream()))
    return uvz


def _plane_layout(t: Optional[torch.Tensor], n_frames: int, n_views: int, out_h: int, out_w: int, channels: int, name: str):
    """mdvt_plane_layout of a (n, H, n_views*W[, C]) side-by-side tensor (or (n, H, W[, C]) for one view)."""
    if t is None:
        return None
    want = (n_frames, out_h, n_views * out_w) + ((channels,) if channels > 1 else ())
    if tuple(t.shape) != want or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous CUDA tensor of shape {want}, got {tuple(t.shape)}")
    es = t.element_size()
    row = n_views * out_w * channels * es
    return _lib.PlaneLayout(t.data_ptr(), out_h * row, out_w * channels * es, row)


def render_views(depth_src: torch.Tensor, colour: torch.Tensor, sources: Sequence[_lib.Source], views: Sequence[Sequence[ViewSpec]],
                 out_w: int, out_h: int, zbuf: torch.Tensor, out_rgb: torch.Tensor, out_mask: Optional[torch.Tensor] = None,
                 out_depth: Optional[torch.Tensor] = None, bg_rgb=(0, 0, 0), fill_rgb=(0, 0, 0), flags: int = 0, near: float = NEAR_PLANE):
    """The generic frame loop in ONE library call: frames (n, H, W, 3) u8 (or (n, H, W) f32 for F32 sources),
    `sources` one per frame or a single one, `views[f]` the cameras of frame f (same count for every frame).
    Views are laid side by side: out_rgb (n, out_h, n_views*out_w, 3) u8, out_mask (n, out_h, n_views*out_w[, 3]) u8,
    out_depth (n, out_h, n_views*out_w) f32.  The z-buffer (n_views or 2*n_views, out_h, out_w) must be empty and is left
    empty; with two sets of planes the odd frames run on the library's second stream next to the even ones."""
    n = depth_src.shape[0]
    packed = isinstance(views, np.ndarray)   # (n, n_views, 16) float32 from pack_views + SOURCE_DTYPE records from pack_sources
    if packed:
        if views.dtype != np.float32 or views.ndim != 3 or views.shape[0] != n or views.shape[2] != 16 or not isinstance(sources, np.ndarray) \
                or sources.dtype != SOURCE_DTYPE:
            raise ValueError("packed cameras: views (n, n_views, 16) float32 and sources as SOURCE_DTYPE records")
        n_views = views.shape[1]
    else:
        n_views = len(views[0])
        if len(views) != n or any(len(v) != n_views for v in views):
            raise ValueError("views must hold the same number of cameras for each of the n frames")
    if len(sources) not in (1, n):
        raise ValueError("sources must hold one entry or one per frame")
    _need(colour, torch.uint8, "colour")
    _need(zbuf, torch.int64, "zbuf")
    if tuple(zbuf.shape) not in ((n_views, out_h, out_w), (2 * n_views, out_h, out_w)):
        raise ValueError(f"zbuf shape {tuple(zbuf.shape)} != ({n_views} or {2 * n_views}, {out_h}, {out_w})")
    zbuf_sets = zbuf.shape[0] // n_views   # two sets: odd frames run on the library's second stream next to the even ones
    if packed:
        src_w, src_h = int(sources["width"][0]), int(sources["height"][0])
        f32_src = int(sources["decoder"][0]) == _lib.SOURCE_F32
        _need(depth_src, torch.float32 if f32_src else torch.uint8, "depth_src")
        if tuple(depth_src.shape[1:3]) != (src_h, src_w) or not depth_src[0].is_contiguous():
            raise ValueError(f"depth_src frames must be dense ({src_h}, {src_w}) planes")
    else:
        for k, src in enumerate(sources):
            _need_source(depth_src[k], src)
        src_w, src_h = sources[0].width, sources[0].height
    if colour.shape[0] != n or colour.shape[1] * colour.shape[2] != src_w * src_h:
        raise ValueError("colour must hold one (H, W, 3) frame per depth frame")
    mask_ch = 3 if flags & FLAG_MASK_RGB else 1
    rgb_l = _plane_layout(_need(out_rgb, torch.uint8, "out_rgb"), n, n_views, out_h, out_w, 3, "out_rgb")
    mask_l = _plane_layout(None if out_mask is None else _need(out_mask, torch.uint8, "out_mask"), n, n_views, out_h, out_w, mask_ch, "out_mask")
    depth_l = _plane_layout(None if out_depth is None else _need(out_depth, torch.float32, "out_depth"), n, n_views, out_h, out_w, 1, "out_depth")
    if packed:
        views_c, sources_c = np.ascontiguousarray(views), np.ascontiguousarray(sources)
        src_arr = (_lib.Source * len(sources_c)).from_buffer(sources_c)
        view_arr = (_lib.View * (n * n_views)).from_buffer(views_c)
    else:
        src_arr = (_lib.Source * len(sources))(*sources)
        view_arr = (_lib.View * (n * n_views))(*[v.to_c() for fv in views for v in fv])
    opt = lambda l: None if l is None else C.byref(l)  # noqa: E731
    _lib.check(_lib.load().mdvt_render_views(_ptr(depth_src), depth_src.stride(0) * depth_src.element_size(), _ptr(colour),
                                             colour.stride(0), n, src_arr, int(len(sources) == n and n > 1), view_arr, n_views,
                                             float(np.float32(near)), int(out_w), int(out_h), _ptr(zbuf), int(zbuf_sets), pack_rgb(bg_rgb), pack_rgb(fill_rgb),
                                             flags, C.byref(rgb_l), opt(mask_l), opt(depth_l), _stream()))
    return out_rgb, out_mask, out_depth


def novel_view_frames(depth_src: torch.Tensor, colour: torch.Tensor, centroid_source: _lib.Source, source: _lib.Source, K: np.ndarray,
                      cam_pos, target, poses, zbuf: torch.Tensor, out_rgb: torch.Tensor, out_mask: Optional[torch.Tensor] = None,
                      bg_rgb=(255, 255, 255), fill_rgb=(255, 255, 255), flags: int = 0, near: float = NEAR_PLANE,
                      sums: Optional[torch.Tensor] = None, views_dev: Optional[torch.Tensor] = None, touched: Optional[torch.Tensor] = None):
    """`3d_view_depthfile.py --render`'s frame loop (:133-255) for a chunk in ONE library call and no host round trip:
    per frame vertex centroid -> look-at camera (on the device) -> splat -> resolve.  `cam_pos`: --x --y --z;
    `target`: three values or None per axis (None: centroid); `poses`: (n, 4, 4) float64 or None.
    Returns (out_rgb, out_mask, sums (n, 4+scratch) f64 device, views (n, 16) f32 device: M 3x4 then fx fy cx cy)."""
    n, h, w = depth_src.shape[0], source.height, source.width
    if n == 0:
        return out_rgb, out_mask, sums, views_dev
    _need_source(depth_src[0], source)
    if not depth_src.is_contiguous():
        raise ValueError("depth_src must be contiguous")
    _need(colour, torch.uint8, "colour")
    _need(zbuf, torch.int64, "zbuf")
    if tuple(zbuf.shape) not in ((1, h, w), (2, h, w)):
        raise ValueError(f"zbuf shape {tuple(zbuf.shape)} != (1 or 2, {h}, {w})")
    zbuf_sets = zbuf.shape[0]   # two planes: odd frames run on the library's second stream next to the even ones
    if colour.shape[0] != n or colour.shape[1] * colour.shape[2] != w * h:
        raise ValueError("colour must hold one (H, W, 3) frame per depth frame")
    dev = depth_src.device
    stride = 4 + _lib.REDUCE_SCRATCH_DOUBLES
    if sums is None:
        sums = torch.empty((n, stride), dtype=torch.float64, device=dev)
    if views_dev is None:
        views_dev = torch.empty((n, 16), dtype=torch.float32, device=dev)
    if tuple(_need(sums, torch.float64, "sums").shape) != (n, stride) or tuple(_need(views_dev, torch.float32, "views_dev").shape) != (n, 16):
        raise ValueError("sums must be (n, 4 + scratch) float64 and views_dev (n, 16) float32")
    need = zbuf_sets * int(_lib.load().mdvt_touched_bytes(w, h))
    if touched is None:
        touched = torch.empty(need, dtype=torch.uint8, device=dev)
    if _need(touched, torch.uint8, "touched").numel() < need:
        raise ValueError(f"touched must hold {need} bytes")
    K = np.asarray(K, dtype=np.float64)
    look = _lib.LookAt()
    cam32 = np.array(cam_pos).astype(np.float32)          # 3d_view_depthfile.py:240
    for a in range(3):
        look.cam_pos[a] = float(cam32[a])
        look.target_set[a] = int(target[a] is not None)
        look.target[a] = 0.0 if target[a] is None else float(target[a])
    look.y_scale = float(K[1, 1] / K[0, 0])
    look.fx = look.fy = float(np.float32(K[0, 0]))
    look.cx, look.cy = float(np.float32(K[0, 2])), float(np.float32(K[1, 2]))
    pose_arr = None
    if poses is not None:
        pm = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(n, 16))
        pose_arr = pm.ctypes.data_as(C.POINTER(C.c_double))
    rgb_l = _plane_layout(_need(out_rgb, torch.uint8, "out_rgb"), n, 1, h, w, 3, "out_rgb")
    mask_ch = 3 if flags & FLAG_MASK_RGB else 1
    mask_l = _plane_layout(None if out_mask is None else _need(out_mask, torch.uint8, "out_mask"), n, 1, h, w, mask_ch, "out_mask")
    _lib.check(_lib.load().mdvt_novel_view_frames(_ptr(depth_src), depth_src.stride(0) * depth_src.element_size(), _ptr(colour), colour.stride(0),
                                                  n, C.byref(centroid_source), C.byref(source), _k4(K), pose_arr, C.byref(look),
                                                  float(np.float32(near)), w, h, _ptr(zbuf), int(zbuf_sets), _ptr(sums), _ptr(views_dev), _ptr(touched), pack_rgb(bg_rgb),
                                                  pack_rgb(fill_rgb), flags, C.byref(rgb_l), None if mask_l is None else C.byref(mask_l),
                                                  _stream()))
    return out_rgb, out_mask, sums, views_dev


def splat_points(xyz: torch.Tensor, views: Sequence[ViewSpec], out_w: int, out_h: int, zbuf: torch.Tensor,
                 near: float = NEAR_PLANE, id_offset: int = 0):
    """Explicit (N, 3) float32 points through the same visibility rule (the reference's point painter)."""
    _need(xyz, torch.float32, "xyz")
    _need(zbuf, torch.int64, "zbuf")
    if xyz.dim() != 2 or xyz.shape[1] != 3:
        raise ValueError("xyz must be (N, 3)")
    if tuple(zbuf.shape) != (len(views), out_h, out_w):
        raise ValueError(f"zbuf shape {tuple(zbuf.shape)} != ({len(views)}, {out_h}, {out_w})")
    arr = (_lib.View * len(views))(*[v.to_c() for v in views])
    _lib.check(_lib.load().mdvt_splat_points(_ptr(xyz), xyz.shape[0], arr, len(views), float(np.float32(near)), int(out_w), int(out_h),
                                             int(id_offset), _ptr(zbuf), _stream()))


# ---------------------------------------------------------------------------------------------
# per-frame reductions
# ---------------------------------------------------------------------------------------------
def _reduce_buffer(device) -> torch.Tensor:
    return torch.empty(4 + _lib.REDUCE_SCRATCH_DOUBLES, dtype=torch.float64, device=device)


def centroid_sums(depth_src: torch.Tensor, source: _lib.Source, K: np.ndarray, pose=None, out: Optional[torch.Tensor] = None):
    """Device tensor [sum X, sum Y, sum Z, n] (float64) of the unprojected (+posed) vertices: Open3D
    get_center() without materialising the vertices (3d_view_depthfile.py:231).  Asynchronous."""
    _need_source(depth_src, source)
    buf = _reduce_buffer(depth_src.device) if out is None else _need(out, torch.float64, "out")
    _lib.check(_lib.load().mdvt_centroid(_ptr(depth_src), C.byref(source), _k4(K), _pose12(pose, C.c_double), _ptr(buf), _stream()))
    return buf[:4]


def depth_sums(depth_src: torch.Tensor, max_depth=100, decoder: str = "D3", bit16: bool = True, mask: Optional[torch.Tensor] = None,
               mask_gt: int = 240, out: Optional[torch.Tensor] = None):
    """Device tensor [sum, count, sum of squares, 0] (float64) of the decoded depth over pixels whose mask byte
    is > mask_gt (find_convergence_depth.py:56-80).  Asynchronous."""
    if decoder == "F32":
        _need(depth_src, torch.float32, "depth_src")
        n = depth_src.numel()
    else:
        _need(depth_src, torch.uint8, "depth_src")
        n = depth_src.numel() // 3
    if mask is not None:
        _need(mask, torch.uint8, "mask")
        if mask.numel() != n:
            raise ValueError("mask must have one byte per pixel")
    buf = _reduce_buffer(depth_src.device) if out is None else _need(out, torch.float64, "out")
    dc = 1.0 if decoder == "F32" else dec_const(max_depth, decoder)
    _lib.check(_lib.load().mdvt_depth_sum(_ptr(depth_src), n, DECODERS[decoder], int(bool(bit16)), dc, _ptr(mask), int(mask_gt), _ptr(buf),
                                          _stream()))
    return buf[:4]


def resolve(zbuf_view: torch.Tensor, colour: torch.Tensor, bg_rgb=(0, 0, 0), fill_rgb=(0, 0, 0), flags: int = 0,
            out_rgb: Optional[torch.Tensor] = None, out_mask: Optional[torch.Tensor] = None, want_depth: bool = False,
            want_ids: bool = False, want_mask: bool = True, out_depth: Optional[torch.Tensor] = None):
    """K3 for one view.  `out_rgb` (H, W, 3) / `out_mask` (H, W[, 3]) may be column slices of wider
    side-by-side tensors (row stride is passed through as the pitch)."""
    _need(zbuf_view, torch.int64, "zbuf_view")
    _need(colour, torch.uint8, "colour")
    out_h, out_w = zbuf_view.shape
    dev = zbuf_view.device
    mask_bpp = 3 if flags & FLAG_MASK_RGB else 1
    if out_rgb is None:
        out_rgb = torch.empty((out_h, out_w, 3), dtype=torch.uint8, device=dev)
    if out_mask is None and want_mask:
        out_mask = torch.empty((out_h, out_w) + ((3,) if mask_bpp == 3 else ()), dtype=torch.uint8, device=dev)
    for t, bpp, name in ((out_rgb, 3, "out_rgb"), (out_mask, mask_bpp, "out_mask")):
        if t is None:
            continue
        if t.dtype != torch.uint8 or not t.is_cuda or t.shape[0] != out_h or t.shape[1] != out_w:
            raise ValueError(f"{name} has the wrong dtype / device / shape")
        inner_ok = (t.dim() == 2 and bpp == 1 and t.stride(1) == 1) or (t.dim() == 3 and t.shape[2] == bpp and t.stride(2) == 1 and t.stride(1) == bpp)
        if not inner_ok:
            raise ValueError(f"{name} rows must be dense")
    depth = out_depth
    if depth is None and want_depth:
        depth = torch.empty((out_h, out_w), dtype=torch.float32, device=dev)
    if depth is not None and (depth.dtype != torch.float32 or not depth.is_cuda or tuple(depth.shape) != (out_h, out_w) or depth.stride(1) != 1):
        raise ValueError("out_depth must be a (H, W) float32 CUDA view with dense rows")
    ids = torch.empty((out_h, out_w), dtype=torch.int32, device=dev) if want_ids else None
    _lib.check(_lib.load().mdvt_resolve(_ptr(zbuf_view), _ptr(colour), out_w, out_h, pack_rgb(bg_rgb), pack_rgb(fill_rgb), flags,
                                        _ptr(out_rgb), out_rgb.stride(0), _ptr(out_mask), 0 if out_mask is None else out_mask.stride(0),
                                        _ptr(depth), 0 if depth is None else depth.stride(0), _ptr(ids),
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
                want_mask: bool = True, out_depth: Optional[torch.Tensor] = None):
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
    if out_depth is not None:
        _need(out_depth, torch.float32, "out_depth")
        if out_depth.numel() != n * h * 2 * w:
            raise ValueError("out_depth must be (n, H, 2W) float32")
    _lib.check(_lib.load().mdvt_stereo_rows(_ptr(depth_rgb), _ptr(colour), n, w, h, _ptr(frames), int(frames.shape[0] == n and n > 1),
                                            pack_rgb(bg_rgb), pack_rgb(fill_rgb), flags, _ptr(out_sbs), _ptr(out_mask), _ptr(out_depth),
                                            _stream()))
    return out_sbs, out_mask


# ---------------------------------------------------------------------------------------------
# normals-coded infill mask: edge vertices, edge-point splat, mask / image painting
# ---------------------------------------------------------------------------------------------
EDGE_ANGLE_DEG = 89.0  # depth_map_tools.py:1192


def edge_vertices(depth_src: torch.Tensor, source: _lib.Source, K: np.ndarray, want_normals: bool = True,
                  angle_threshold_deg: float = EDGE_ANGLE_DEG, flags: Optional[torch.Tensor] = None,
                  normals: Optional[torch.Tensor] = None):
    """Edge test of the reference's mesh builder (depth_map_tools.py:1243-1376).  Returns (flags (H, W) u8 with 1 at
    vertices of too-oblique triangles, normals (H, W, 3) float64 -- defined at flagged vertices only -- or None)."""
    _need_source(depth_src, source)
    h, w = source.height, source.width
    dev = depth_src.device
    if flags is None:
        flags = torch.empty((h, w), dtype=torch.uint8, device=dev)
    if normals is None and want_normals:
        normals = torch.empty((h, w, 3), dtype=torch.float64, device=dev)
    scratch = torch.empty(((h - 1) * (w - 1),), dtype=torch.uint8, device=dev)
    _lib.check(_lib.load().mdvt_edge_vertices(_ptr(depth_src), C.byref(source), _k4(K), float(angle_threshold_deg), _ptr(scratch),
                                              _ptr(_need(flags, torch.uint8, "flags")), _ptr(normals), _stream()))
    return flags, normals


def edge_vertices_xyz(xyz: torch.Tensor, height: int, width: int, want_normals: bool = True, angle_threshold_deg: float = EDGE_ANGLE_DEG):
    """The same edge test on explicit grid-organised vertices (H*W, 3) float64 (create_mesh_from_point_cloud's input)."""
    _need(xyz, torch.float64, "xyz")
    if xyz.numel() != height * width * 3:
        raise ValueError(f"xyz must hold {height}x{width} points")
    dev = xyz.device
    flags = torch.empty((height, width), dtype=torch.uint8, device=dev)
    normals = torch.empty((height, width, 3), dtype=torch.float64, device=dev) if want_normals else None
    scratch = torch.empty(((height - 1) * (width - 1),), dtype=torch.uint8, device=dev)
    _lib.check(_lib.load().mdvt_edge_vertices_xyz(_ptr(xyz), int(width), int(height), float(angle_threshold_deg), _ptr(scratch), _ptr(flags),
                                                  _ptr(normals), _stream()))
    return flags, normals


def edge_splat(depth_src: torch.Tensor, source: _lib.Source, K: np.ndarray, flags: torch.Tensor, pose, K_render: np.ndarray,
               out_w: int, out_h: int, zbuf: torch.Tensor):
    """Flagged vertices -> edge points -> eye camera (`pose`, 4x4 float64) -> z-buffer (out_h, out_w) int64."""
    _need_source(depth_src, source)
    _need(flags, torch.uint8, "flags")
    _need(zbuf, torch.int64, "zbuf")
    if tuple(zbuf.shape) != (out_h, out_w):
        raise ValueError(f"zbuf shape {tuple(zbuf.shape)} != ({out_h}, {out_w})")
    k32 = np.asarray(K_render).astype(np.float32).astype(np.float64)  # the reference hands cv2 a float32 camera matrix
    _lib.check(_lib.load().mdvt_edge_splat(_ptr(depth_src), C.byref(source), _k4(K), _ptr(flags), _pose12(pose, C.c_double), _k4(k32),
                                           int(out_w), int(out_h), _ptr(zbuf), _stream()))


def edge_resolve(zbuf: torch.Tensor, depth_src: torch.Tensor, source: _lib.Source, K: np.ndarray, normals: Optional[torch.Tensor], pose,
                 colour: torch.Tensor, hole_mask: torch.Tensor, mask_img: torch.Tensor, image: Optional[torch.Tensor] = None,
                 bg_rgb=(0, 255, 0), code_normals: bool = True):
    """Per target pixel: the (pre-inpainting) mask image and the edge colours painted into `image`.  hole_mask (H, W)
    u8, mask_img / image (H, W, 3) u8 may be column slices of side-by-side tensors.  Leaves `zbuf` empty."""
    _need(zbuf, torch.int64, "zbuf")
    _need(colour, torch.uint8, "colour")
    out_h, out_w = zbuf.shape
    for t, ch, name in ((hole_mask, 1, "hole_mask"), (mask_img, 3, "mask_img"), (image, 3, "image")):
        if t is None:
            continue
        ok = t.dtype == torch.uint8 and t.is_cuda and t.shape[0] == out_h and t.shape[1] == out_w and t.stride(-1) == 1
        ok = ok and ((ch == 1 and t.dim() == 2) or (ch == 3 and t.dim() == 3 and t.shape[2] == 3 and t.stride(1) == 3))
        if not ok:
            raise ValueError(f"{name} must be a ({out_h}, {out_w}{', 3' if ch == 3 else ''}) u8 CUDA view with dense rows")
    _lib.check(_lib.load().mdvt_edge_resolve(_ptr(zbuf), _ptr(depth_src), C.byref(source), _k4(K), _ptr(normals), _pose12(pose, C.c_double),
                                             _ptr(colour), _ptr(hole_mask), hole_mask.stride(0), out_w, out_h, pack_rgb(bg_rgb),
                                             int(bool(code_normals)), _ptr(image), 0 if image is None else image.stride(0), _ptr(mask_img),
                                             mask_img.stride(0), _stream()))


def normal_march_infill(image: torch.Tensor, hole_mask: torch.Tensor, mask_img: torch.Tensor, max_steps: int = 400) -> torch.Tensor:
    """stereo_rerender.infill_using_normals (:155-240) for one eye, in place on `image` (H, W, 3) u8; hole_mask (H, W) u8,
    mask_img (H, W, 3) u8 = the final (inpainted, blurred) mask image.  All three may be column slices."""
    h, w = hole_mask.shape
    for t, ch, name in ((image, 3, "image"), (hole_mask, 1, "hole_mask"), (mask_img, 3, "mask_img")):
        ok = t.dtype == torch.uint8 and t.is_cuda and t.shape[0] == h and t.shape[1] == w and t.stride(-1) == 1
        ok = ok and ((ch == 1 and t.dim() == 2) or (ch == 3 and t.dim() == 3 and t.shape[2] == 3 and t.stride(1) == 3))
        if not ok:
            raise ValueError(f"{name} must be a ({h}, {w}{', 3' if ch == 3 else ''}) u8 CUDA view with dense rows")
    _lib.check(_lib.load().mdvt_normal_march_infill(_ptr(image), image.stride(0), _ptr(hole_mask), hole_mask.stride(0), _ptr(mask_img),
                                                    mask_img.stride(0), w, h, int(max_steps), _stream()))
    return image


def pack_mask_bits(mask: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(..., W) u8 hole mask {0, 255} -> (..., W / 8) u8, one bit per pixel, most significant bit first (np.packbits order).
    W must be a multiple of 8 so that rows stay byte aligned."""
    _need(mask, torch.uint8, "mask")
    if mask.shape[-1] % 8:
        raise ValueError("the mask width must be a multiple of 8")
    shape = tuple(mask.shape[:-1]) + (mask.shape[-1] // 8,)
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=mask.device)
    if tuple(_need(out, torch.uint8, "out").shape) != shape:
        raise ValueError(f"out must be {shape}")
    _lib.check(_lib.load().mdvt_pack_mask_bits(_ptr(mask), mask.numel(), _ptr(out), _stream()))
    return out


def unpack_mask_bits(bits, width: Optional[int] = None) -> np.ndarray:
    """Host side of pack_mask_bits: (..., W / 8) u8 (NumPy array or CPU tensor) -> (..., W) u8 {0, 255}."""
    arr = bits.numpy() if isinstance(bits, torch.Tensor) else np.asarray(bits)
    out = np.unpackbits(arr, axis=-1)
    if width is not None:
        out = out[..., :width]
    return out * np.uint8(255)


def normal_march_infill_f32(image: torch.Tensor, hole_mask: torch.Tensor, normal_map: torch.Tensor, max_steps: int = 400) -> torch.Tensor:
    """stereo_rerender.infill_using_normals(color_img, hole_mask, normal_map, max_steps) (:155-240) in place on `image`
    (H, W, 3) u8; hole_mask (H, W) u8 (non-zero = hole); normal_map (H, W, 3) float32, dense."""
    h, w = hole_mask.shape
    for t, ch, name in ((image, 3, "image"), (hole_mask, 1, "hole_mask")):
        ok = t.dtype == torch.uint8 and t.is_cuda and t.shape[0] == h and t.shape[1] == w and t.stride(-1) == 1
        ok = ok and ((ch == 1 and t.dim() == 2) or (ch == 3 and t.dim() == 3 and t.shape[2] == 3 and t.stride(1) == 3))
        if not ok:
            raise ValueError(f"{name} must be a ({h}, {w}{', 3' if ch == 3 else ''}) u8 CUDA view with dense rows")
    _need(normal_map, torch.float32, "normal_map")
    if tuple(normal_map.shape) != (h, w, 3):
        raise ValueError(f"normal_map must be ({h}, {w}, 3)")
    _lib.check(_lib.load().mdvt_normal_march_infill_f32(_ptr(image), image.stride(0), _ptr(hole_mask), hole_mask.stride(0), _ptr(normal_map),
                                                        w, h, int(max_steps), _stream()))
    return image


def calculate_normals(depth: torch.Tensor, K) -> torch.Tensor:
    """depth_map_tools.calculate_normals (:20-60): (H, W) float32 depth -> (H, W, 3) float32 unit normals."""
    _need(depth, torch.float32, "depth")
    if depth.dim() != 2:
        raise ValueError("depth must be (H, W)")
    h, w = depth.shape
    out = torch.empty((h, w, 3), dtype=torch.float32, device=depth.device)
    _lib.check(_lib.load().mdvt_calculate_normals(_ptr(depth), w, h, _k4(K), _ptr(out), _stream()))
    return out


# ---------------------------------------------------------------------------------------------
# stereo with a convergence rotation: fused target-row kernel
# ---------------------------------------------------------------------------------------------
def conv_frames(sources: Sequence[_lib.Source], views: Sequence[Sequence[ViewSpec]], near: float = NEAR_PLANE) -> np.ndarray:
    """Host array of mdvt_conv_frame (as (n, 40) float32) from per-frame sources and [left, right] cameras; raises if a
    camera is not `rotation about y + shift along x` (the only poses the fused kernel handles)."""
    n = len(views)
    if len(sources) not in (1, n):
        raise ValueError("sources must hold one entry or one per frame")
    arr = (_lib.ConvFrame * n)()
    for k in range(n):
        s = sources[k if len(sources) == n else 0]
        if s.decoder != _lib.DECODE_D1 or not s.bit16 or s.grid_sx != 1.0 or s.grid_sy != 1.0:
            raise ValueError("the fused convergence kernel takes 16-bit D1 wire-format frames on the exact pixel grid")
        f = arr[k]
        f.dec_const, f.depth_scale, f.near_plane = s.dec_const, s.depth_scale, float(np.float32(near))
        f.fx, f.fy, f.cx, f.cy = s.fx, s.fy, s.cx, s.cy
        if len(views[k]) != 2:
            raise ValueError("two cameras (left, right) per frame")
        for e, v in enumerate(views[k]):
            c = v.to_c()
            m = [c.M[i] for i in range(12)]
            if any(m[i] != 0.0 for i in (1, 4, 6, 7, 9, 11)) or m[5] != 1.0:
                raise ValueError("camera pose is not a y-rotation plus an x-shift")
            f.view[e] = c
    return np.frombuffer(bytes(arr), dtype=np.float32).reshape(n, C.sizeof(_lib.ConvFrame) // 4).copy()


def conv_frames_packed(sources: np.ndarray, views: np.ndarray, near: float = NEAR_PLANE) -> np.ndarray:
    """conv_frames from packed cameras (pack_sources records, pack_views rows (n, 2, 16)): (n, 40) float32, vectorised."""
    n = views.shape[0]
    if views.shape[1:] != (2, 16) or len(sources) not in (1, n):
        raise ValueError("two cameras (left, right) per frame; one source or one per frame")
    src = sources if len(sources) == n else np.repeat(sources, n)
    if np.any(src["decoder"] != _lib.DECODE_D1) or np.any(src["bit16"] == 0) or np.any(src["grid_sx"] != 1.0) or np.any(src["grid_sy"] != 1.0):
        raise ValueError("the fused convergence kernel takes 16-bit D1 wire-format frames on the exact pixel grid")
    if np.any(views[:, :, [1, 4, 6, 7, 9, 11]] != 0.0) or np.any(views[:, :, 5] != 1.0):
        raise ValueError("camera pose is not a y-rotation plus an x-shift")
    out = np.zeros((n, C.sizeof(_lib.ConvFrame) // 4), dtype=np.float32)
    out[:, 0], out[:, 1], out[:, 2] = src["dec_const"], src["depth_scale"], np.float32(near)
    out[:, 4], out[:, 5], out[:, 6], out[:, 7] = src["fx"], src["fy"], src["cx"], src["cy"]
    out[:, 8:] = views.reshape(n, 32)
    return out


def conv_vrows_supported(frames_host: np.ndarray, width: int, height: int) -> bool:
    """True when every frame of the HOST array `frames_host` ((n, 40) float32 rows of mdvt_conv_frame) is a pose the
    virtual-row kernel handles at this size (mdvt_stereo_conv_vrows_supported: no CUDA call)."""
    arr = np.ascontiguousarray(frames_host, dtype=np.float32)
    if arr.ndim != 2 or arr.shape[1] != C.sizeof(_lib.ConvFrame) // 4:
        raise ValueError("frames_host must be (n, 40) float32 rows of mdvt_conv_frame")
    return bool(_lib.load().mdvt_stereo_conv_vrows_supported(arr.ctypes.data, arr.shape[0], int(width), int(height)))


def stereo_conv_rows(depth_rgb: torch.Tensor, colour: torch.Tensor, frames: torch.Tensor, bg_rgb=(0, 0, 0), fill_rgb=(0, 0, 0),
                     flags: int = 0, out_sbs: Optional[torch.Tensor] = None, out_mask: Optional[torch.Tensor] = None,
                     want_mask: bool = True, out_depth: Optional[torch.Tensor] = None, kernel: str = "rows",
                     status: Optional[torch.Tensor] = None):
    """Fused convergence-stereo kernel over a batch: depth_rgb / colour (n, H, W, 3) u8, frames (n, 40) float32 CUDA
    tensor from conv_frames().  Same outputs as stereo_rows; bit-identical to project_splat + resolve.
    kernel "vrows": the virtual-source-row kernel (mdvt_stereo_conv_vrows; poses must pass conv_vrows_supported), `status`
    an optional zeroed (n,) int32 CUDA tensor that receives 1 for frames whose geometry left the kernel's limits;
    kernel "rows": the target-row kernel with per-column source-row prediction (mdvt_stereo_conv_rows, any width)."""
    if kernel not in ("rows", "vrows"):
        raise ValueError("kernel is 'rows' or 'vrows'")
    _need(depth_rgb, torch.uint8, "depth_rgb")
    _need(colour, torch.uint8, "colour")
    _need(frames, torch.float32, "frames")
    if depth_rgb.dim() != 4 or depth_rgb.shape[-1] != 3 or depth_rgb.shape != colour.shape:
        raise ValueError("depth_rgb and colour must both be (n, H, W, 3)")
    n, h, w, _ = depth_rgb.shape
    if tuple(frames.shape) != (n, C.sizeof(_lib.ConvFrame) // 4):
        raise ValueError("frames must be (n, 40) float32 rows of mdvt_conv_frame")
    dev = depth_rgb.device
    mask_bpp = 3 if flags & FLAG_MASK_RGB else 1
    if out_sbs is None:
        out_sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device=dev)
    _need(out_sbs, torch.uint8, "out_sbs")
    if want_mask and out_mask is None:
        out_mask = torch.empty((n, h, 2 * w) + ((3,) if mask_bpp == 3 else ()), dtype=torch.uint8, device=dev)
    if out_mask is not None:
        _need(out_mask, torch.uint8, "out_mask")
        if out_mask.numel() != n * h * 2 * w * mask_bpp:
            raise ValueError("out_mask has the wrong size")
    if out_sbs.numel() != n * h * 2 * w * 3:
        raise ValueError("out_sbs has the wrong size")
    if out_depth is not None:
        _need(out_depth, torch.float32, "out_depth")
        if out_depth.numel() != n * h * 2 * w:
            raise ValueError("out_depth must be (n, H, 2W) float32")
    if kernel == "vrows":
        if status is not None:
            _need(status, torch.int32, "status")
            if status.numel() != n:
                raise ValueError("status must hold one int32 per frame")
        _lib.check(_lib.load().mdvt_stereo_conv_vrows(_ptr(depth_rgb), _ptr(colour), n, w, h, _ptr(frames), pack_rgb(bg_rgb), pack_rgb(fill_rgb),
                                                      flags, _ptr(out_sbs), _ptr(out_mask), _ptr(out_depth), _ptr(status), _stream()))
        return out_sbs, out_mask
    _lib.check(_lib.load().mdvt_stereo_conv_rows(_ptr(depth_rgb), _ptr(colour), n, w, h, _ptr(frames), pack_rgb(bg_rgb), pack_rgb(fill_rgb),
                                                 flags, _ptr(out_sbs), _ptr(out_mask), _ptr(out_depth), _stream()))
    return out_sbs, out_mask
