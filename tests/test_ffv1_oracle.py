"""CPU: oracle/ffv1_oracle.py pinned against libavcodec's FFV1 as this image's OpenCV wheel drives it (the codec of
every result video of the reference: stereo_rerender.py:420-442,941; depth_frames_helper.py:125-161).

* the configuration record written by the oracle is byte-identical to the CodecPrivate of an OpenCV-written file;
* key and non-key frame packets of an OpenCV-written file decode to the original frames;
* whole packets encoded by the oracle (key frame and the non-key frames after it) are byte-identical to libavcodec's;
* all-key-frame streams with many slices written by the oracle + mkv_join are decoded bit-exactly by OpenCV.
"""
import os

import cv2
import numpy as np
import pytest

from metric_depth_video_toolbox_b200 import mkv_join
from oracle import ffv1_oracle as fo

W, H = 64, 48
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ffv1_opencv.npz")


def _golden():
    g = np.load(GOLDEN)
    bounds = np.concatenate([[0], np.cumsum(g["packet_sizes"])])
    packets = [g["packets"][bounds[k]:bounds[k + 1]].tobytes() for k in range(len(g["packet_sizes"]))]
    return g["frames_bgr"], g["config"].tobytes(), packets, [bool(k) for k in g["keys"]]


def test_committed_golden_clip():
    """tests/golden/ffv1_opencv.npz (oracle/make_ffv1_golden.py: cv2.VideoWriter output of OpenCV 4.13.0 / avcodec
    62.11.100): the oracle decodes libavcodec's packets to the frames and re-encodes them byte for byte, whatever the
    OpenCV build of the machine the tests run on writes."""
    frames, config, packets, keys = _golden()
    cfg = fo.parse_config(config)
    assert fo.crc32_mpeg(config) == 0 and fo.write_config(cfg, cfg["num_h_slices"], cfg["num_v_slices"]) == config
    assert (cfg["version"], cfg["ac"], cfg["colorspace"], cfg["transparency"], cfg["num_h_slices"], cfg["num_v_slices"]) == (3, 0, 1, 1, 2, 2)
    assert keys == [True, False, False, False, False]
    h, w = frames.shape[1:3]
    dec = [fo.SliceState(cfg) for _ in range(4)]
    enc = [fo.SliceState(cfg) for _ in range(4)]
    for k, f in enumerate(frames):
        out, key = fo.decode_frame(packets[k], cfg, w, h, dec)
        assert bool(key) == keys[k] and np.array_equal(out[..., :3], f) and (out[..., 3] == 255).all()
        assert fo.encode_frame(_bgra(f), cfg, keys[k], enc) == packets[k], f"packet {k}"


def _frames(n, seed=0):
    rng = np.random.default_rng(seed)
    out = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(n)]
    if n > 2:
        out[2][:] = out[2][:1, :1]                       # a flat frame: run mode all the way
    if n > 3:
        out[3] = cv2.GaussianBlur(out[3], (0, 0), 3)     # smooth content: small residuals, adaptive k
    return out


def _bgra(f):
    return np.dstack([f, np.full(f.shape[:2], 255, np.uint8)])


@pytest.fixture(scope="module")
def cv_file(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("ffv1") / "cv.mkv")
    frames = _frames(5)
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 24.0, (W, H))
    if not wr.isOpened():
        pytest.skip("this OpenCV build has no FFV1 encoder")
    for f in frames:
        wr.write(f)
    wr.release()
    pk = mkv_join.MkvPackets(path)
    if pk.codec_private() != _golden()[1]:
        pytest.skip("this OpenCV / libavcodec build writes other FFV1 stream parameters than the committed golden clip's; "
                    "the byte-identity pins run against tests/golden/ffv1_opencv.npz only")
    return frames, pk, fo.parse_config(pk.codec_private())


def test_stream_facts(cv_file):
    _, pk, cfg = cv_file
    assert (cfg["version"], cfg["micro_version"], cfg["ac"], cfg["colorspace"], cfg["bits"]) == (3, 4, 0, 1, 8)
    assert cfg["transparency"] == 1 and cfg["ec"] == 1
    assert (cfg["num_h_slices"], cfg["num_v_slices"]) == (2, 2)
    assert cfg["context_count"][0] == 666
    assert [p[2] for p in pk.packets] == [True, False, False, False, False]


def test_config_record_is_byte_identical(cv_file):
    _, pk, cfg = cv_file
    assert fo.write_config(cfg, 2, 2) == pk.codec_private()
    assert fo.crc32_mpeg(pk.codec_private()) == 0


def test_decodes_opencv_packets(cv_file):
    frames, pk, cfg = cv_file
    ss = [fo.SliceState(cfg) for _ in range(4)]
    for k, f in enumerate(frames):
        out, key = fo.decode_frame(pk.payload(k), cfg, W, H, ss)
        assert key == (1 if k == 0 else 0)
        assert np.array_equal(out[..., :3], f) and (out[..., 3] == 255).all()


def test_encoded_packets_are_byte_identical(cv_file):
    frames, pk, cfg = cv_file
    ss = [fo.SliceState(cfg) for _ in range(4)]
    for k, f in enumerate(frames):
        assert fo.encode_frame(_bgra(f), cfg, k == 0, ss) == pk.payload(k), f"packet {k}"


@pytest.mark.parametrize("nh,nv", [(2, 2), (8, 8), (16, 12), (5, 7)])
def test_many_slice_streams_decode_in_opencv(cv_file, tmp_path, nh, nv):
    frames, pk, base = cv_file
    cfg = dict(base, num_h_slices=nh, num_v_slices=nv)
    extra = fo.write_config(base, nh, nv)
    assert fo.parse_config(extra)["num_v_slices"] == nv
    tracks = mkv_join.replace_codec_private(pk.tracks, extra)
    packets = []
    for f in frames:
        ss = [fo.SliceState(cfg) for _ in range(nh * nv)]
        packets.append((fo.encode_frame(_bgra(f), cfg, True, ss), True))
    path = str(tmp_path / "many.mkv")
    assert mkv_join.write_stream(path, pk.ebml_header, tracks, packets, len(packets), 24.0) == len(frames)
    cap = cv2.VideoCapture(path)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == len(frames)
    for f in frames:
        ok, got = cap.read()
        assert ok and np.array_equal(got, f)
    assert not cap.read()[0]
    assert not os.path.exists(path + ".joining")
