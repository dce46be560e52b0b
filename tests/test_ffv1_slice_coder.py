"""CPU: the FFV1 stream constants the C-ABI library writes on the host (mdvt_ffv1_stream_setup: configuration record,
slice headers) and the per-slice coder the device runs (csrc/mdvt_ffv1_slice.h, compiled here as plain C++ by
tests/support/ffv1_slice_host.cpp) against libavcodec (through OpenCV) and oracle/ffv1_oracle.py.  No device call."""
import ctypes as C
import os

import cv2
import numpy as np
import pytest

from metric_depth_video_toolbox_b200 import ffv1_gpu, mkv_join
from oracle import ffv1_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_coder(tmp_path_factory):
    from ffv1_host import HostCoder

    return HostCoder(str(tmp_path_factory.mktemp("ffv1_host")))


def _cv_file(path, frames, fps=24.0):
    h, w = frames[0].shape[:2]
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), fps, (w, h))
    if not wr.isOpened():
        pytest.skip("this OpenCV build has no FFV1 encoder")
    for f in frames:
        wr.write(f)
    wr.release()
    pk = mkv_join.MkvPackets(path)
    golden = np.load(os.path.join(ROOT, "tests", "golden", "ffv1_opencv.npz"))["config"].tobytes()
    if pk.codec_private() != golden:
        pytest.skip("this OpenCV / libavcodec build writes other FFV1 stream parameters than the committed golden clip's "
                    "(tests/golden/ffv1_opencv.npz); the byte-identity pins against libavcodec run on the golden clip only")
    return pk


def _content(w, h, seed=0):
    rng = np.random.default_rng(seed)
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    flat = np.empty_like(noise)
    flat[:] = noise[:1, :1]
    smooth = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 4)
    yy, xx = np.mgrid[0:h, 0:w]
    ramp = np.dstack([(xx * 3 + yy) & 255, (xx + yy * 2) & 255, (xx ^ yy) & 255]).astype(np.uint8)
    ramp[h // 3: h // 2, w // 4: w // 2] = (0, 255, 0)      # the green / black mask look: long runs with hard edges
    extremes = np.where(rng.random((h, w, 1)) < 0.5, 0, 255).astype(np.uint8).repeat(3, 2)
    extremes[..., 1] = 255 - extremes[..., 0]               # worst-case residuals: escape codes
    return [noise, flat, smooth, ramp, extremes]


def test_config_record_and_headers_match_libavcodec(tmp_path):
    """2 x 2 slices with alpha is what OpenCV writes: the record must be its CodecPrivate byte for byte, and the slice
    headers must be the leading bytes of its key-frame slices."""
    w, h = 64, 48
    frames = _content(w, h)[:1]
    pk = _cv_file(str(tmp_path / "cv.mkv"), frames)
    config, headers, lens = ffv1_gpu.stream_setup(w, h, 2, 2, alpha=True)
    assert config == pk.codec_private()
    cfg = fo.parse_config(config)
    packet = pk.payload(0)
    for (a, b), head, n in zip(fo.slice_ranges(packet, 4, cfg["ec"]), headers, lens):
        assert packet[a:a + n] == head[:n].tobytes()
    for nh, nv, alpha in ((32, 32, False), (59, 17, False), (5, 7, True)):
        config, headers, lens = ffv1_gpu.stream_setup(3840, 1080, nh, nv, alpha)
        c = fo.parse_config(config)
        assert (c["num_h_slices"], c["num_v_slices"], c["transparency"], c["ec"], c["version"], c["ac"]) == (nh, nv, int(alpha), 1, 3, 0)
        assert c["quant_tables"] == cfg["quant_tables"] and fo.crc32_mpeg(config) == 0
        assert lens.min() >= 2 and lens.max() <= 16 and headers.shape == (nh * nv, 16)


def test_golden_clip_pins_setup_coder_and_decoder(host_coder):
    """Against the committed libavcodec output (tests/golden/ffv1_opencv.npz, independent of the local OpenCV build): the
    configuration record of mdvt_ffv1_stream_setup(2 x 2, alpha) is its CodecPrivate, the slice coder reproduces its key
    frame packet, the slice decoder reads that packet and refuses the non-key frames after it."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ffv1_opencv.npz"))
    frames, config = g["frames_bgr"], g["config"].tobytes()
    bounds = np.concatenate([[0], np.cumsum(g["packet_sizes"])])
    packets = [g["packets"][bounds[k]:bounds[k + 1]].tobytes() for k in range(len(bounds) - 1)]
    h, w = frames.shape[1:3]
    assert ffv1_gpu.stream_setup(w, h, 2, 2, True, 0)[0] == config
    assert host_coder(frames[0], 2, 2, True, True) == packets[0]
    assert host_coder(frames[0][..., ::-1], 2, 2, True, False) == packets[0]
    rc, out = host_coder.decode(packets[0], w, h, 2, 2, True, True)
    assert rc == 0 and np.array_equal(out, frames[0])
    assert host_coder.decode(packets[1], w, h, 2, 2, True, True)[0] == -1
    for k in range(1, len(frames)):     # every frame as a key frame of its own: what the device writer produces
        rc, out = host_coder.decode(host_coder(frames[k], 2, 2, True, True), w, h, 2, 2, True, True)
        assert rc == 0 and np.array_equal(out, frames[k])


def test_packet_equals_libavcodec_key_frame(host_coder, tmp_path):
    """Same parameters as OpenCV (2 x 2 slices, alpha plane): whole key-frame packets are byte-identical."""
    w, h = 96, 40
    for k, f in enumerate(_content(w, h, seed=3)):
        pk = _cv_file(str(tmp_path / f"cv{k}.mkv"), [f])
        assert host_coder(f, 2, 2, True, True) == pk.payload(0), f"content {k}"
        rgb = cv2.cvtColor(f, cv2.COLOR_BGR2RGB)
        assert host_coder(rgb, 2, 2, True, False) == pk.payload(0), f"content {k} (RGB order)"


@pytest.mark.parametrize("w,h,nh,nv,alpha,model", [(64, 48, 8, 8, False, 0), (64, 48, 16, 12, True, 0), (70, 33, 5, 7, False, 0),
                                                   (33, 17, 33, 17, False, 0), (48, 32, 1, 1, False, 0), (64, 48, 8, 8, False, 1),
                                                   (70, 33, 5, 7, True, 1), (48, 32, 1, 1, False, 1), (64, 48, 8, 8, False, 2),
                                                   (70, 33, 5, 7, True, 2), (48, 32, 1, 1, False, 2)])
def test_packet_equals_oracle(host_coder, w, h, nh, nv, alpha, model):
    """The oracle codes with whatever quant tables the configuration record carries: model 0 = libavcodec's (666
    contexts), model 1 = the 5-level table (63 contexts), model 2 = the 3-level table (14 contexts)."""
    base = fo.parse_config(ffv1_gpu.stream_setup(w, h, nh, nv, alpha, model)[0])
    assert base["context_count"][0] == (666, 63, 14)[model] and base["quant_table_count"] == (1 if model else 2)
    for k, f in enumerate(_content(w, h, seed=5)):
        ss = [fo.SliceState(base) for _ in range(nh * nv)]
        want = fo.encode_frame(np.dstack([f, np.full((h, w), 255, np.uint8)]), base, True, ss)
        assert host_coder(f, nh, nv, alpha, True, model) == want, f"content {k}"


@pytest.mark.parametrize("w,h,nh,nv,alpha,model", [(256, 144, 16, 9, False, 0), (256, 144, 32, 32, False, 0), (200, 120, 7, 5, True, 0),
                                                   (256, 144, 16, 9, False, 1), (200, 120, 7, 5, True, 1), (256, 144, 16, 9, False, 2),
                                                   (200, 120, 7, 5, True, 2)])
def test_stream_decodes_in_opencv(host_coder, tmp_path, w, h, nh, nv, alpha, model):
    """Packets of the slice coder + this library's configuration record + mkv_join's muxer -> OpenCV returns the frames."""
    frames = _content(w, h, seed=9)
    header, tracks = ffv1_gpu.container_template(w, h, 24.0)
    config = ffv1_gpu.stream_setup(w, h, nh, nv, alpha, model)[0]
    path = str(tmp_path / "gpu_style.mkv")
    mux = mkv_join.StreamWriter(path, header, mkv_join.replace_codec_private(tracks, config), 24.0)
    for f in frames:
        mux.add(host_coder(f, nh, nv, alpha, True, model), True)
    assert mux.close() == len(frames)
    cap = cv2.VideoCapture(path)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == len(frames)
    assert (int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))) == (w, h)
    for k, f in enumerate(frames):
        ok, got = cap.read()
        assert ok and np.array_equal(got, f), f"frame {k}"
    assert not cap.read()[0]


@pytest.mark.parametrize("w,h,nh,nv,alpha,model", [(64, 48, 8, 8, False, 0), (70, 33, 5, 7, True, 0), (33, 17, 33, 17, False, 0),
                                                   (48, 32, 1, 1, False, 0), (256, 144, 16, 9, False, 0), (64, 48, 8, 8, False, 1),
                                                   (70, 33, 5, 7, True, 1), (256, 144, 16, 9, False, 1), (64, 48, 8, 8, False, 2),
                                                   (70, 33, 5, 7, True, 2), (256, 144, 16, 9, False, 2)])
def test_decoder_mirrors_encoder_and_oracle(host_coder, w, h, nh, nv, alpha, model):
    """The slice decoder the device runs (host-stepped): packets of the slice coder and of the oracle's encoder decode
    to the source frames in either channel order; damaged packets are reported, not decoded."""
    base = fo.parse_config(ffv1_gpu.stream_setup(w, h, nh, nv, alpha, model)[0])
    for k, f in enumerate(_content(w, h, seed=11)):
        packet = host_coder(f, nh, nv, alpha, True, model)
        rc, out = host_coder.decode(packet, w, h, nh, nv, alpha, True, model)
        assert rc == 0 and np.array_equal(out, f), f"content {k}"
        if k == 0 and 16 * nh * nv <= w * h <= 64 * 48 * 2:      # the other model's decoder must not reproduce the frame
            assert not np.array_equal(host_coder.decode(packet, w, h, nh, nv, alpha, True, (model + 1) % 3)[1], f)
        rc, out = host_coder.decode(packet, w, h, nh, nv, alpha, False, model)     # RGB-order output of a BGR-order source
        assert rc == 0 and np.array_equal(out, f[..., ::-1]), f"content {k} (channel order)"
        if w * h <= 64 * 48:
            ss = [fo.SliceState(base) for _ in range(nh * nv)]
            want = fo.encode_frame(np.dstack([f, np.full((h, w), 255, np.uint8)]), base, True, ss)
            rc, out = host_coder.decode(want, w, h, nh, nv, alpha, True, model)
            assert rc == 0 and np.array_equal(out, f), f"content {k} (oracle packet)"
    bad = bytearray(packet)
    bad[0] ^= 0x40                                         # slice 0's header (key-frame bit / slice position)
    assert host_coder.decode(bytes(bad), w, h, nh, nv, alpha, True, model)[0] == -1
    bad = bytearray(packet)
    bad[len(bad) // 2] ^= 0x10                             # anywhere else: the slice's CRC catches it before decoding
    assert host_coder.decode(bytes(bad), w, h, nh, nv, alpha, True, model)[0] in (-4, -10)
    assert host_coder.decode(packet[:-1], w, h, nh, nv, alpha, True, model)[0] != 0       # truncated: sizes do not add up
    assert host_coder.decode(packet + b"\0", w, h, nh, nv, alpha, True, model)[0] != 0


def test_decoder_rejects_opencv_non_key_frames(host_coder, tmp_path):
    """OpenCV's own packets: the key frame has this coder's parameters (2 x 2, alpha) and decodes; the frames after it
    carry coder state over and start with a different slice header -> refused (they stay with cv2.VideoCapture)."""
    w, h = 96, 40
    frames = _content(w, h, seed=2)[:3]
    pk = _cv_file(str(tmp_path / "cv.mkv"), frames)
    rc, out = host_coder.decode(pk.payload(0), w, h, 2, 2, True, True)
    assert rc == 0 and np.array_equal(out, frames[0])
    assert host_coder.decode(pk.payload(1), w, h, 2, 2, True, True)[0] == -1


def test_parse_config_accepts_only_this_librarys_streams(tmp_path):
    from metric_depth_video_toolbox_b200 import _lib

    lib = _lib.load()
    nh, nv, alpha, model = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    for grid in ((59, 17, 0, 0), (2, 2, 1, 0), (1, 1, 0, 0), (59, 17, 0, 1), (4, 3, 1, 1), (59, 17, 0, 2), (4, 3, 1, 2)):
        cfg = ffv1_gpu.stream_setup(3840, 1080, grid[0], grid[1], bool(grid[2]), grid[3])[0]
        assert lib.mdvt_ffv1_parse_config(cfg, len(cfg), 3840, 1080, C.byref(nh), C.byref(nv), C.byref(alpha), C.byref(model)) == 0
        assert (nh.value, nv.value, alpha.value, model.value) == grid
    pk = _cv_file(str(tmp_path / "cv.mkv"), _content(64, 48)[:1])
    cfg = pk.codec_private()                # OpenCV's record is byte-identical to this library's 2 x 2 + alpha record
    assert lib.mdvt_ffv1_parse_config(cfg, len(cfg), 64, 48, C.byref(nh), C.byref(nv), C.byref(alpha), C.byref(model)) == 0
    assert (nh.value, nv.value, alpha.value, model.value) == (2, 2, 1, 0)
    broken = bytearray(cfg)
    broken[20] ^= 1                         # inside the quant tables
    assert lib.mdvt_ffv1_parse_config(bytes(broken), len(broken), 64, 48, C.byref(nh), C.byref(nv), C.byref(alpha), C.byref(model)) == -2
    assert lib.mdvt_ffv1_parse_config(b"\x00\x01\x02\x03", 4, 64, 48, C.byref(nh), C.byref(nv), C.byref(alpha), C.byref(model)) == -2


def test_rank_segments_join_at_packet_level(host_coder, tmp_path):
    """What `stereo_rerender --gpu_ffv1` does under torchrun: every rank leaves a segment .mkv + `<segment>.plan.json`
    (GpuFfv1Writer(join_on_close=False)); rank 0 stitches them with video_io.join_plans.  Segments are made here with the
    host-stepped coder; an empty rank (fewer GOP-aligned ranges than ranks) leaves an empty plan."""
    import json

    from metric_depth_video_toolbox_b200 import video_io

    w, h, nh, nv = 128, 72, 4, 3
    frames = _content(w, h, seed=13) + _content(w, h, seed=14)[:2]
    header, tracks = ffv1_gpu.container_template(w, h, 25.0)
    tracks = mkv_join.replace_codec_private(tracks, ffv1_gpu.stream_setup(w, h, nh, nv, False)[0])
    parts = []
    for rank, (a, b) in enumerate(((0, 4), (4, 7), (7, 7))):
        seg = str(tmp_path / f"out.mkv.rank{rank:02d}.mkv")
        mux = mkv_join.StreamWriter(seg, header, tracks, 25.0)
        for f in frames[a:b]:
            mux.add(host_coder(f, nh, nv, False, True), True)
        n = mux.close()
        json.dump({"fps": 25.0, "plan": [(seg, n)] if n else []}, open(seg + ".plan.json", "w"))
        parts.append(seg)
    out = str(tmp_path / "out.mkv")
    assert video_io.join_plans([video_io.load_plan(p) for p in parts], out, 25.0) == len(frames)
    got = video_io.read_clip(out, rgb=False)
    assert got.shape[0] == len(frames) and all(np.array_equal(g, f) for g, f in zip(got, frames))
    cap = cv2.VideoCapture(out)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == len(frames) and abs(cap.get(cv2.CAP_PROP_FPS) - 25.0) < 1e-6
    cap.set(cv2.CAP_PROP_POS_FRAMES, 5)                      # every frame is a key frame: random access is exact
    ok, f5 = cap.read()
    assert ok and np.array_equal(f5, frames[5])
    assert not any(os.path.exists(p + ".plan.json") for p in parts)
    assert not os.path.exists(parts[0]) and not os.path.exists(parts[1])   # joined lanes are removed


def test_round_trip_property(host_coder):
    """Random sizes, slice grids, context models, channel orders and content classes: decode(encode(x)) == x, and the
    packet is the oracle's for the small ones."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    @settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(w=st.integers(1, 96), h=st.integers(1, 40), fx=st.integers(1, 12), fy=st.integers(1, 12), alpha=st.booleans(), model=st.integers(0, 2),
           bgr=st.booleans(), kind=st.integers(0, 4), seed=st.integers(0, 2**31 - 1))
    def check(w, h, fx, fy, alpha, model, bgr, kind, seed):
        nh, nv = min(fx, w), min(fy, h)
        rng = np.random.default_rng(seed)
        if kind == 0:
            f = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        elif kind == 1:
            f = np.full((h, w, 3), rng.integers(0, 256, 3), np.uint8)
        elif kind == 2:
            f = (rng.integers(0, 2, (h, w, 1), dtype=np.uint8) * 255).repeat(3, 2)
            f[..., 1] = 255 - f[..., 1]
        elif kind == 3:
            f = (128 + rng.integers(-3, 4, (h, w, 3))).astype(np.uint8)
        else:
            yy, xx = np.mgrid[0:h, 0:w]
            f = np.dstack([(3 * xx + yy) & 255, (xx + 2 * yy) & 255, (xx * yy) & 255]).astype(np.uint8)
        packet = host_coder(f, nh, nv, alpha, bgr, model)
        rc, out = host_coder.decode(packet, w, h, nh, nv, alpha, bgr, model)
        assert rc == 0 and np.array_equal(out, f)
        if w * h <= 600:
            cfg = fo.parse_config(ffv1_gpu.stream_setup(w, h, nh, nv, alpha, model)[0])
            ss = [fo.SliceState(cfg) for _ in range(nh * nv)]
            src = f if bgr else f[..., ::-1]
            assert fo.encode_frame(np.dstack([src, np.full((h, w), 255, np.uint8)]), cfg, True, ss) == packet

    check()


def test_slice_coder_under_address_sanitizer(tmp_path):
    """The coder text the device runs, built for the host with ASan + UBSan: round trips over 1-pixel / ragged slices and
    decoding of bit-flipped packets never leave the packet, the state block or the frame."""
    import subprocess

    from metric_depth_video_toolbox_b200 import build

    exe = str(tmp_path / "ffv1_sanitize")
    sup = os.path.join(ROOT, "tests", "support")
    proc = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I",
                           os.path.join(ROOT, "metric_depth_video_toolbox_b200", "csrc"), os.path.join(sup, "ffv1_sanitize_main.cpp"),
                           os.path.join(sup, "ffv1_slice_host.cpp"), "-ldl", "-o", exe], capture_output=True, text=True)
    if proc.returncode != 0:
        pytest.skip("no sanitizer runtime for g++ here: " + proc.stderr[-200:])
    run = subprocess.run([exe, build.build()], capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert run.returncode == 0 and "done, 0 bad" in run.stdout, run.stdout[-2000:] + run.stderr[-4000:]


def test_slice_grid():
    for w, h in ((3840, 1080), (1920, 1080), (3840, 2160), (640, 480), (64, 48), (7, 3)):
        nh, nv = ffv1_gpu.slice_grid(w, h)
        assert 1 <= nh <= w and 1 <= nv <= h and nh * nv <= 1024
    assert ffv1_gpu.slice_grid(3840, 1080)[0] * ffv1_gpu.slice_grid(3840, 1080)[1] > 900
    assert ffv1_gpu.slice_grid(64, 48) == (1, 1)
