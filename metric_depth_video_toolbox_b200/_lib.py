"""ctypes binding of libmdvt_b200.so (include/mdvt_b200.h).

There is no CPU fallback: if the shared library is missing the loader tries to build it with
nvcc, and raises if that fails; every compute entry point raises `MdvtError` on a non-zero
status (e.g. no CUDA device).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

OK = 0
DECODE_D1, DECODE_D2, DECODE_D3, SOURCE_F32 = 0, 1, 2, 3
DECODERS = {"D1": DECODE_D1, "D2": DECODE_D2, "D3": DECODE_D3, "F32": SOURCE_F32}
REDUCE_SCRATCH_DOUBLES = 4096
FLAG_BG_COLLIDE, FLAG_RESET_ZBUF, FLAG_MASK_RGB, FLAG_ANYWIDTH = 0x1, 0x2, 0x4, 0x8
ZBUF_EMPTY = 0xFFFFFFFFFFFFFFFF
ABI_VERSION = 11


class MdvtError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libmdvt_b200 status {status}: {message}")
        self.status = status


class Source(C.Structure):
    """mdvt_source"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("decoder", C.c_int32), ("bit16", C.c_int32),
                ("dec_const", C.c_float), ("depth_scale", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("grid_sx", C.c_float), ("grid_sy", C.c_float)]


class View(C.Structure):
    """mdvt_view"""
    _fields_ = [("M", C.c_float * 12), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class ConvFrame(C.Structure):
    """mdvt_conv_frame (40 x float32; built on the host, copied to the device as bytes)"""
    _fields_ = [("dec_const", C.c_float), ("depth_scale", C.c_float), ("near_plane", C.c_float), ("reserved", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("view", View * 2)]


class PlaneLayout(C.Structure):
    """mdvt_plane_layout (strides in bytes)"""
    _fields_ = [("base", C.c_void_p), ("frame_stride", C.c_int64), ("view_stride", C.c_int64), ("row_pitch", C.c_int64)]


class LookAt(C.Structure):
    """mdvt_lookat"""
    _fields_ = [("cam_pos", C.c_double * 3), ("target", C.c_double * 3), ("target_set", C.c_int32 * 3), ("reserved", C.c_int32),
                ("y_scale", C.c_double), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class StereoFrame(C.Structure):
    """mdvt_stereo_frame (4 x float32; built as a (n, 4) float32 tensor on the device)"""
    _fields_ = [("dec_const", C.c_float), ("depth_scale", C.c_float), ("fx_half_ipd", C.c_float),
                ("near_plane", C.c_float)]


_u8p, _u32p, _u64p, _i32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p  # device pointers travel as integers
_f32p, _f64p, _stream = C.c_void_p, C.c_void_p, C.c_void_p

_PROTOTYPES = {
    "mdvt_abi_version": (C.c_int, []),
    "mdvt_version": (C.c_char_p, []),
    "mdvt_last_error": (C.c_char_p, []),
    "mdvt_device_info": (C.c_int, [C.POINTER(C.c_int)] * 5),
    "mdvt_decode_depth": (C.c_int, [_u8p, C.c_int64, C.c_int, C.c_int, C.c_float, _u32p, _f32p, _stream]),
    "mdvt_encode_depth": (C.c_int, [_f32p, C.c_int64, C.c_double, C.c_int, C.c_int, _u32p, _u8p, _stream]),
    "mdvt_encode_depth_f64": (C.c_int, [_f64p, C.c_int64, C.c_double, C.c_int, C.c_int, _u32p, _u8p, _stream]),
    "mdvt_unproject_f32": (C.c_int, [_u8p, C.POINTER(Source), C.POINTER(C.c_float), _f32p, _stream]),
    "mdvt_unproject_f64": (C.c_int, [_u8p, C.POINTER(Source), C.POINTER(C.c_double), C.POINTER(C.c_double), _f64p, _stream]),
    "mdvt_depth_to_grey": (C.c_int, [_u8p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, _stream]),
    "mdvt_pack_mask_bits": (C.c_int, [_u8p, C.c_int64, _u8p, _stream]),
    "mdvt_touchly_depth": (C.c_int, [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_int, _u8p, C.c_int64, _stream]),
    "mdvt_remap_bilinear_u8x3": (C.c_int, [_u8p, C.c_int, C.c_int, C.c_int64, _f32p, _f32p, C.c_int, C.c_int, C.c_uint32, _u8p, C.c_int64,
                                           _stream]),
    "mdvt_transform_points_f64": (C.c_int, [_f64p, C.c_int64, C.POINTER(C.c_double), _f64p, _stream]),
    "mdvt_project_points_f64": (C.c_int, [_f64p, C.c_int64, C.POINTER(C.c_double), _f64p, _stream]),
    "mdvt_zbuf_clear": (C.c_int, [_u64p, C.c_int64, _stream]),
    "mdvt_codes_to_depth": (C.c_int, [_u32p, C.c_int64, C.c_int, C.c_float, _f32p, _stream]),
    "mdvt_codes_to_pixels": (C.c_int, [_u32p, C.c_int64, C.c_int, C.c_int, _u8p, _stream]),
    "mdvt_project_splat": (C.c_int, [_u8p, C.POINTER(Source), C.POINTER(View), C.c_int, C.c_float, C.c_int, C.c_int,
                                     C.c_uint32, _u64p, _f32p, _stream]),
    "mdvt_splat_points": (C.c_int, [_f32p, C.c_int64, C.POINTER(View), C.c_int, C.c_float, C.c_int, C.c_int, C.c_uint32,
                                    _u64p, _stream]),
    "mdvt_centroid": (C.c_int, [_u8p, C.POINTER(Source), C.POINTER(C.c_double), C.POINTER(C.c_double), _f64p, _stream]),
    "mdvt_depth_sum": (C.c_int, [_u8p, C.c_int64, C.c_int, C.c_int, C.c_float, _u8p, C.c_int, _f64p, _stream]),
    "mdvt_resolve": (C.c_int, [_u64p, _u8p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, C.c_int64,
                               _u8p, C.c_int64, _f32p, C.c_int64, _i32p, _stream]),
    "mdvt_render_views": (C.c_int, [_u8p, C.c_int64, _u8p, C.c_int64, C.c_int, C.POINTER(Source), C.c_int, C.POINTER(View), C.c_int,
                                    C.c_float, C.c_int, C.c_int, _u64p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(PlaneLayout),
                                    C.POINTER(PlaneLayout), C.POINTER(PlaneLayout), _stream]),
    "mdvt_touched_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "mdvt_novel_view_frames": (C.c_int, [_u8p, C.c_int64, _u8p, C.c_int64, C.c_int, C.POINTER(Source), C.POINTER(Source), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.POINTER(LookAt), C.c_float, C.c_int, C.c_int, _u64p, C.c_int, _f64p, C.c_void_p, _u8p,
                                         C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(PlaneLayout), C.POINTER(PlaneLayout), _stream]),
    "mdvt_edge_vertices": (C.c_int, [_u8p, C.POINTER(Source), C.POINTER(C.c_double), C.c_double, _u8p, _u8p, _f64p, _stream]),
    "mdvt_edge_vertices_xyz": (C.c_int, [_f64p, C.c_int, C.c_int, C.c_double, _u8p, _u8p, _f64p, _stream]),
    "mdvt_edge_splat": (C.c_int, [_u8p, C.POINTER(Source), C.POINTER(C.c_double), _u8p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.c_int, C.c_int, _u64p, _stream]),
    "mdvt_edge_resolve": (C.c_int, [_u64p, _u8p, C.POINTER(Source), C.POINTER(C.c_double), _f64p, C.POINTER(C.c_double), _u8p, _u8p,
                                    C.c_int64, C.c_int, C.c_int, C.c_uint32, C.c_int, _u8p, C.c_int64, _u8p, C.c_int64, _stream]),
    "mdvt_normal_march_infill": (C.c_int, [_u8p, C.c_int64, _u8p, C.c_int64, _u8p, C.c_int64, C.c_int, C.c_int, C.c_int, _stream]),
    "mdvt_normal_march_infill_f32": (C.c_int, [_u8p, C.c_int64, _u8p, C.c_int64, _f32p, C.c_int, C.c_int, C.c_int, _stream]),
    "mdvt_calculate_normals": (C.c_int, [_f32p, C.c_int, C.c_int, C.POINTER(C.c_double), _f32p, _stream]),
    "mdvt_stereo_conv_rows": (C.c_int, [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, _u8p,
                                        _f32p, _stream]),
    "mdvt_stereo_conv_vrows": (C.c_int, [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, _u8p,
                                         _f32p, _i32p, _stream]),
    "mdvt_stereo_conv_vrows_supported": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "mdvt_stereo_rows": (C.c_int, [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32,
                                   C.c_uint32, _u8p, _u8p, _f32p, _stream]),
    "mdvt_ffv1_stream_setup": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                         C.c_void_p, C.c_void_p]),
    "mdvt_ffv1_slice_capacity": (C.c_int64, [C.c_int] * 5),
    "mdvt_ffv1_state_bytes": (C.c_int64, [C.c_int] * 5),
    "mdvt_ffv1_encode_frames": (C.c_int, [_u8p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          _u8p, _i32p, C.c_void_p, _u8p, C.c_int64, _i32p, C.c_void_p, _u8p, _stream]),
    "mdvt_ffv1_parse_config": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int)]),
    "mdvt_ffv1_decode_frames": (C.c_int, [_u8p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p,
                                          _i32p, C.c_void_p, C.c_void_p, _u8p, C.c_int64, C.c_int64, _i32p, _stream]),
}

_lock = threading.Lock()
_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the .so is absent or stale and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("MDVT_B200_LIB") or _build.LIB_PATH   # MDVT_B200_LIB: a differently built copy of the library (kernel tuning experiments)
        if not os.path.exists(path):
            _build.build()  # raises if nvcc is missing: there is no other implementation to fall back to
        lib = C.CDLL(path)
        for name, (restype, argtypes) in _PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError here == header and library disagree
            fn.restype, fn.argtypes = restype, argtypes
        if lib.mdvt_abi_version() != ABI_VERSION:
            raise ImportError(f"{path}: ABI {lib.mdvt_abi_version()} != binding {ABI_VERSION}; rebuild the library")
        _lib = lib
    return _lib


def check(status: int):
    if status != OK:
        raise MdvtError(status, load().mdvt_last_error().decode("utf-8", "replace"))


def exported_names():
    return list(_PROTOTYPES)
