"""CPU: the C-ABI library builds, loads, exports every symbol include/mdvt_b200.h declares, and its
argument validation (which runs before any CUDA call) reports errors the documented way.
No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from metric_depth_video_toolbox_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mdvt_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MDVT_API\s+[\w\s\*]+?\b(mdvt_\w+)\s*\(", text)))


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.isfile(path)
    lib = _lib.load()
    assert lib.mdvt_abi_version() == _lib.ABI_VERSION
    assert b"sm_100a" in lib.mdvt_version()


def test_every_declared_symbol_is_exported_and_bound():
    declared = declared_symbols()
    assert len(declared) >= 12
    out = subprocess.run(["nm", "-D", "--defined-only", build.build()], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    stray = sorted(s for s in exported if s.startswith("mdvt_") and s not in declared)
    assert not stray, f"exported but not declared: {stray}"
    assert sorted(_lib.exported_names()) == declared  # the Python binding covers the whole header


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Source) == 4 * 4 + 8 * 4
    assert C.sizeof(_lib.View) == 16 * 4
    assert C.sizeof(_lib.StereoFrame) == 16
    assert C.sizeof(_lib.PlaneLayout) == 32
    assert C.sizeof(_lib.ConvFrame) == 160
    assert C.sizeof(_lib.LookAt) == 88


def test_sass_contains_tma_and_atomics():
    """The built library really holds the Blackwell paths the design names: 1-D TMA bulk copies
    (UBLKCP), mbarrier transactions (SYNCS), shared-memory and 64-bit global atomic min."""
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True)
    if sass.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    for mnemonic in ("UBLKCP", "SYNCS", "ATOMS", "MIN.64"):
        assert mnemonic in sass.stdout, mnemonic
    assert "sm_100a" in sass.stdout
    # --fmad=false: the parity contract forbids contraction in the geometry kernels
    body = sass.stdout.split("stereo_rows_w32_kernel")[1]
    assert "FMUL" in body and "ATOMS.MIN" in body and "UBLKCP" in body


def test_argument_validation_without_a_device():
    lib = _lib.load()
    assert lib.mdvt_decode_depth(None, -1, 0, 1, 1.0, None, None, None) == -1
    assert b"negative" in lib.mdvt_last_error()
    assert lib.mdvt_decode_depth(None, 16, 7, 1, 1.0, None, None, None) == -1  # unknown decoder
    assert lib.mdvt_decode_depth(None, 16, _lib.DECODE_D2, 0, 1.0, None, None, None) == -2  # 24-bit is D1-only
    assert lib.mdvt_decode_depth(None, 0, 0, 1, 1.0, None, None, None) == 0  # empty input is a no-op
    assert lib.mdvt_encode_depth(None, 4, 0.0, 1, 1, None, None, None) == -1
    assert lib.mdvt_stereo_rows(None, None, 1, 70000, 4, None, 0, 0, 0, 0, None, None, None, None) == -2
    assert b"65535" in lib.mdvt_last_error()
    assert lib.mdvt_stereo_rows(None, None, 0, 64, 4, None, 0, 0, 0, 0, None, None, None, None) == 0
    assert lib.mdvt_zbuf_clear(None, 0, None) == 0
    src = _lib.Source()
    assert lib.mdvt_project_splat(None, C.byref(src), None, 1, 1e-4, 4, 4, 0, None, None, None) == -1  # 0x0 frame
    assert lib.mdvt_splat_points(None, -1, None, 1, 1e-4, 4, 4, 0, None, None) == -1
    assert lib.mdvt_splat_points(None, 0, C.byref(_lib.View()), 1, 1e-4, 4, 4, 0, None, None) == 0
    assert lib.mdvt_codes_to_depth(None, 0, 0, 1.0, None, None) == 0
    assert lib.mdvt_codes_to_pixels(None, -3, 1, 1, None, None) == -1
    assert lib.mdvt_transform_points_f64(None, 0, (C.c_double * 12)(), None, None) == 0
    assert lib.mdvt_project_points_f64(None, 5, None, None, None) == -1
    assert lib.mdvt_depth_sum(None, 16, 9, 1, 1.0, None, 240, None, None) == -1  # unknown decoder
    assert lib.mdvt_centroid(None, C.byref(src), None, None, None, None) == -1
    assert lib.mdvt_render_views(None, 0, None, 0, 0, None, 0, None, 2, 1e-4, 4, 4, None, 1, 0, 0, 0, None, None, None, None) == 0
    assert lib.mdvt_render_views(None, 0, None, 0, 1, None, 0, None, 2, 1e-4, 4, 4, None, 3, 0, 0, 0, None, None, None, None) == -1  # zbuf_sets
    assert lib.mdvt_render_views(None, 0, None, 0, 1, None, 0, None, 9, 1e-4, 4, 4, None, 1, 0, 0, 0, None, None, None, None) == -1
    assert lib.mdvt_resolve(None, None, 4, 4, 0, 0, 0, None, 0, None, 0, None, 0, None, None) == -1
    assert lib.mdvt_stereo_conv_rows(None, None, 1, 5000, 4, None, 0, 0, 0, None, None, None, None) == -2  # column does not fit 12 bits
    assert lib.mdvt_stereo_conv_rows(None, None, 0, 64, 4, None, 0, 0, 0, None, None, None, None) == 0
    assert lib.mdvt_stereo_conv_vrows(None, None, 1, 70, 4, None, 0, 0, 0, None, None, None, None, None) == -2  # width not a multiple of 32
    assert b"multiples of 32" in lib.mdvt_last_error()
    assert lib.mdvt_stereo_conv_vrows(None, None, 0, 64, 4, None, 0, 0, 0, None, None, None, None, None) == 0
    assert lib.mdvt_stereo_conv_vrows_supported(None, 1, 64, 4) == 0
    assert lib.mdvt_remap_bilinear_u8x3(None, 4, 4, 12, None, None, 4, 4, 0, None, 12, None) == -1
    assert lib.mdvt_normal_march_infill(None, 12, None, 4, None, 12, 4, 4, 400, None) == -1
    assert lib.mdvt_edge_vertices(None, C.byref(src), None, 89.0, None, None, None, None) == -1
    assert lib.mdvt_depth_to_grey(None, 4, 0, 1, 1.0, 1.0, 12, 1, None, None) == -1  # 12-bit output does not exist
    assert lib.mdvt_touchly_depth(None, 4, 4, 0, 1, 1.0, 1.0, 5.0, 5.0, 1.0, 0, None, 12, None) == -1  # max must exceed min


def test_ops_refuse_cpu_tensors():
    import torch

    from metric_depth_video_toolbox_b200 import ops

    with pytest.raises(TypeError, match="no CPU path"):
        ops.decode_depth(torch.zeros((4, 3), dtype=torch.uint8), 100)
    with pytest.raises(TypeError, match="no CPU path"):
        ops.stereo_rows(torch.zeros((1, 2, 16, 3), dtype=torch.uint8), torch.zeros((1, 2, 16, 3), dtype=torch.uint8),
                        torch.zeros((1, 4)))


def test_ffv1_argument_validation_without_a_device():
    lib = _lib.load()
    cfg = (C.c_uint8 * 64)()
    n = C.c_int()
    hdr = (C.c_uint8 * (16 * 1024))()
    lens = (C.c_int32 * 1024)()
    assert lib.mdvt_ffv1_stream_setup(64, 48, 2, 2, 1, 0, cfg, 64, C.byref(n), hdr, lens) == 0 and n.value == 42
    assert lib.mdvt_ffv1_stream_setup(64, 48, 2, 2, 0, 1, cfg, 64, C.byref(n), hdr, lens) == 0 and 0 < n.value < 42   # one quant-table set
    assert lib.mdvt_ffv1_stream_setup(64, 48, 2, 2, 1, 0, cfg, 8, C.byref(n), hdr, lens) == -1                         # record does not fit
    assert lib.mdvt_ffv1_stream_setup(64, 48, 33, 32, 0, 0, cfg, 64, C.byref(n), hdr, lens) == -1
    assert b"1024" in lib.mdvt_last_error()
    assert lib.mdvt_ffv1_stream_setup(64, 48, 65, 1, 0, 0, cfg, 64, C.byref(n), hdr, lens) == -1                       # more slices than columns
    assert lib.mdvt_ffv1_stream_setup(64, 48, 2, 2, 0, 2, cfg, 64, C.byref(n), hdr, lens) == 0 and 0 < n.value < 42   # the 14-context tables
    assert lib.mdvt_ffv1_stream_setup(64, 48, 2, 2, 0, 3, cfg, 64, C.byref(n), hdr, lens) == -1                        # unknown context model
    assert lib.mdvt_ffv1_slice_capacity(3840, 1080, 59, 17, 0) % 16 == 0 and lib.mdvt_ffv1_slice_capacity(3840, 1080, 59, 17, 0) > 66 * 64 * 3
    assert lib.mdvt_ffv1_slice_capacity(0, 1080, 1, 1, 0) == -1
    assert lib.mdvt_ffv1_state_bytes(2, 59, 17, 0, 0) == 2 * 1003 * 2 * 666 * 8
    assert lib.mdvt_ffv1_state_bytes(2, 59, 17, 1, 1) == 2 * 1003 * 3 * 63 * 8
    assert lib.mdvt_ffv1_state_bytes(2, 59, 17, 0, 2) == 2 * 1003 * 2 * 14 * 8
    assert lib.mdvt_ffv1_state_bytes(2, 59, 17, 0, 3) == -1
    assert lib.mdvt_ffv1_encode_frames(None, 0, 0, 0, 64, 48, 2, 2, 0, 0, 0, None, None, None, None, 0, None, None, None, None) == 0  # no frames
    assert lib.mdvt_ffv1_encode_frames(None, 0, 0, 1, 64, 48, 2, 2, 0, 0, 0, None, None, None, None, 16, None, None, None, None) == -1
    assert b"capacity" in lib.mdvt_last_error()
    assert lib.mdvt_ffv1_decode_frames(None, None, 0, 64, 48, 2, 2, 0, 0, 0, None, None, None, None, None, 0, 0, None, None) == 0
    assert lib.mdvt_ffv1_decode_frames(None, None, 1, 64, 48, 2, 2, 0, 0, 0, None, None, None, None, None, 1, 1, None, None) == -1
    assert b"pitch" in lib.mdvt_last_error()
