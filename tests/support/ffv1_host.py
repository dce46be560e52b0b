"""Test support: builds tests/support/ffv1_slice_host.cpp (csrc/mdvt_ffv1_slice.h -- the per-slice FFV1 coder the device
runs -- compiled as plain C++) and wraps it.  Nothing in the package uses this."""
import ctypes as C
import os
import subprocess

import numpy as np

from metric_depth_video_toolbox_b200 import ffv1_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class HostCoder:
    def __init__(self, build_dir: str):
        so = os.path.join(build_dir, "ffv1_slice_host.so")
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "metric_depth_video_toolbox_b200", "csrc"),
                        os.path.join(ROOT, "tests", "support", "ffv1_slice_host.cpp"), "-o", so], check=True)
        lib = self.lib = C.CDLL(so)
        lib.ffv1_host_encode_frame.restype = C.c_longlong
        lib.ffv1_host_encode_frame.argtypes = [C.c_void_p, C.c_longlong] + [C.c_int] * 7 + [C.c_void_p] * 3 + [C.c_longlong]
        lib.ffv1_host_decode_frame.restype = C.c_int
        lib.ffv1_host_decode_frame.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong] + [C.c_int] * 7 + [C.c_void_p] * 2

    def encode(self, frame, nh, nv, alpha, bgr, model=0) -> bytes:
        """One frame (H, W, 3) uint8 -> its packet (all slices, raster order)."""
        h, w = frame.shape[:2]
        frame = np.ascontiguousarray(frame)
        _, headers, lens = ffv1_gpu.stream_setup(w, h, nh, nv, alpha, model)
        cap = w * h * 12 + 4096 * nh * nv
        out = np.zeros(cap, np.uint8)
        n = self.lib.ffv1_host_encode_frame(frame.ctypes.data, frame.strides[0], w, h, nh, nv, 3 + int(alpha), int(bgr), int(model),
                                            headers.ctypes.data, lens.ctypes.data, out.ctypes.data, cap)
        assert n > 0
        return out[:n].tobytes()

    __call__ = encode

    def decode(self, packet, w, h, nh, nv, alpha, bgr, model=0):
        """-> (status, frame (H, W, 3) uint8)."""
        _, headers, lens = ffv1_gpu.stream_setup(w, h, nh, nv, alpha, model)
        buf = np.frombuffer(packet, np.uint8)
        out = np.full((h, w, 3), 0xA5, np.uint8)
        rc = self.lib.ffv1_host_decode_frame(buf.ctypes.data, len(packet), out.ctypes.data, out.strides[0], w, h, nh, nv, 3 + int(alpha),
                                             int(bgr), int(model), headers.ctypes.data, lens.ctypes.data)
        return rc, out
