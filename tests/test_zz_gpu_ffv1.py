"""GPU: the FFV1 encoder (mdvt_ffv1_encode_frames, one device thread per slice) through the C ABI.

* packets are byte-identical to the same slice coder stepped on the host (tests/support/ffv1_slice_host.cpp), which
  tests/test_ffv1_slice_coder.py pins against libavcodec and oracle/ffv1_oracle.py;
* files written by GpuFfv1Writer are read back bit-exactly by OpenCV (libavcodec's decoder), at the 3840x1080 size of a
  1080p side-by-side result too;
* the offsets / sizes the device reports add up.
(Named zz: it runs after every other GPU test.)"""
import os

import cv2
import numpy as np
import pytest
import torch

from metric_depth_video_toolbox_b200 import ffv1_gpu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def content(w, h, n, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        kind = k % 4
        if kind == 0:
            f = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 3)
            f[h // 4: h // 2, w // 8: w // 3] = rng.integers(0, 256, 3)
        elif kind == 1:
            f = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        elif kind == 2:
            f = np.zeros((h, w, 3), np.uint8)
            f[h // 3: 2 * h // 3, w // 5: w // 2] = (0, 255, 0)   # green / black hole mask
        else:
            f = np.where(rng.random((h, w, 1)) < 0.5, 0, 255).astype(np.uint8).repeat(3, 2)
            f[..., 1] = 255 - f[..., 0]
        out.append(f)
    return np.stack(out)


@pytest.fixture(scope="module")
def host_coder(tmp_path_factory):
    from ffv1_host import HostCoder

    return HostCoder(str(tmp_path_factory.mktemp("ffv1_host")))


@pytest.mark.parametrize("w,h,slices,alpha,rgb,model", [(256, 144, (16, 9), False, True, 0), (200, 120, (7, 5), True, False, 0),
                                                        (320, 200, (32, 32), False, False, 0), (64, 48, (1, 1), False, True, 0)])
def test_packets_equal_host_stepped_coder(host_coder, w, h, slices, alpha, rgb, model):
    frames = content(w, h, 6, seed=w)
    enc = ffv1_gpu.Ffv1Encoder(w, h, DEV, max_frames=8, slices=slices, alpha=alpha, context_model=model)
    dev = torch.from_numpy(frames).to(DEV)
    packets = enc.encode(dev, rgb=rgb)
    assert len(packets) == len(frames)
    for k, f in enumerate(frames):
        assert packets[k] == host_coder(f, slices[0], slices[1], alpha, not rgb, model), f"frame {k}"
    s = enc.per_frame
    sizes = enc.sizes[: len(frames) * s].cpu().numpy().astype(np.int64)
    offsets = enc.offsets[: len(frames) * s + 1].cpu().numpy()
    assert offsets[0] == 0 and np.array_equal(np.diff(offsets), sizes)
    assert sizes.max() <= enc.capacity


@pytest.mark.parametrize("w,h,n,model", [(640, 360, 9, 0), (3840, 1080, 3, 0)])
def test_writer_files_decode_bit_exactly_in_opencv(tmp_path, w, h, n, model):
    frames = content(w, h, n, seed=7)      # RGB order, as the renderers hold them
    path = str(tmp_path / "gpu.mkv")
    wr = ffv1_gpu.GpuFfv1Writer(path, 24.0, (w, h), device=DEV, batch=4, context_model=model)
    dev = torch.from_numpy(frames).to(DEV)
    wr.write(dev[:5], rgb=True)
    wr.write(dev[5:], rgb=True)
    assert wr.close() == n and os.path.isfile(path) and not os.path.exists(path + ".joining")
    cap = cv2.VideoCapture(path)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == n
    assert abs(cap.get(cv2.CAP_PROP_FPS) - 24.0) < 1e-6
    for k in range(n):
        ok, got = cap.read()
        assert ok and np.array_equal(cv2.cvtColor(got, cv2.COLOR_BGR2RGB), frames[k]), f"frame {k}"
    assert not cap.read()[0]


def test_rejects_host_tensors_and_bad_sizes():
    enc = ffv1_gpu.Ffv1Encoder(64, 48, DEV, max_frames=2)
    with pytest.raises(TypeError):
        enc.encode(torch.zeros((1, 48, 64, 3), dtype=torch.uint8))
    with pytest.raises(ValueError):
        enc.encode(torch.zeros((3, 48, 64, 3), dtype=torch.uint8, device=DEV))
    with pytest.raises(ValueError):
        enc.encode(torch.zeros((1, 40, 64, 3), dtype=torch.uint8, device=DEV))
    assert enc.encode(torch.zeros((0, 48, 64, 3), dtype=torch.uint8, device=DEV)) == []


def test_cli_stereo_rerender_with_gpu_writer_equals_default_writer(tmp_path, depth_video=True):
    """stereo_rerender (result videos coded on the device: the default) writes the same frames (main, mask and SBS depth
    video) as with the host writers (MDVT_FFV1_WRITER=host).
    Without --create_sbs_depth_video the front end hands the device tensors of the row kernel straight to the coder."""
    import stereo_rerender   # the root-level launcher
    from metric_depth_video_toolbox_b200 import video_io
    from metric_depth_video_toolbox_b200.synth import SyntheticClip

    w, h, n = 192, 108, 7
    depth, colour = SyntheticClip(w, h, n).frames()
    dpath, cpath = str(tmp_path / "depth.mkv"), str(tmp_path / "colour.mkv")
    video_io.write_clip(dpath, depth, 24.0)
    video_io.write_clip(cpath, colour, 24.0)
    argv = ["--depth_video", dpath, "--color_video", cpath, "--xfov", "60", "--infill_mask", "--green_and_black_infill_mask",
            "--dont_place_points_in_edges", "--chunk_frames", "3"] + (["--create_sbs_depth_video"] if depth_video else [])
    names = [dpath + "_stereo.mkv", dpath + "_stereo.mkv_infillmask.mkv"] + ([dpath + "_stereo.mkv_depth.mkv"] if depth_video else [])
    os.environ["MDVT_FFV1_WRITER"] = "host"   # the reference's writers (cv2.VideoWriter lanes); the device coder is the default
    try:
        assert stereo_rerender.main(argv) == 0
    finally:
        os.environ.pop("MDVT_FFV1_WRITER", None)
    want = [video_io.read_clip(p) for p in names]
    assert not mkv_packets_all_key(names[0])
    for p in names:
        os.remove(p)
    assert stereo_rerender.main(argv) == 0     # default: coded on the device
    assert mkv_packets_all_key(names[0])
    for p, ref in zip(names, want):
        got = video_io.read_clip(p)
        assert ref.shape[0] == n and got.shape == ref.shape and np.array_equal(got, ref), p
        cap = cv2.VideoCapture(p)
        assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == n
    assert not [f for f in os.listdir(tmp_path) if "_tmp_" in f or f.endswith(".joining")]


@pytest.mark.parametrize("w,h,slices,alpha,rgb,model", [(256, 144, (16, 9), False, True, 0), (200, 120, (7, 5), True, False, 0),
                                                        (320, 200, (32, 32), False, True, 0)])
def test_device_decoder_mirrors_encoder(w, h, slices, alpha, rgb, model):
    frames = content(w, h, 6, seed=w + 1)
    enc = ffv1_gpu.Ffv1Encoder(w, h, DEV, max_frames=8, slices=slices, alpha=alpha, context_model=model)
    packets = enc.encode(torch.from_numpy(frames).to(DEV), rgb=rgb)
    dec = ffv1_gpu.Ffv1Decoder.for_config(enc.config, w, h, DEV, max_frames=8)
    assert (dec.nh, dec.nv, dec.alpha, dec.context_model) == (slices[0], slices[1], alpha, model)
    got = dec.decode(packets, rgb=rgb)
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), frames)
    swapped = dec.decode(packets[:2], rgb=not rgb)
    assert np.array_equal(swapped.cpu().numpy(), frames[:2, ..., ::-1])
    bad = bytearray(packets[1])
    bad[len(bad) // 2] ^= 0xFF      # damaged bits: the frame still decodes to *something* or overruns, the others are intact
    try:
        out = dec.decode([packets[0], bytes(bad), packets[2]], rgb=rgb).cpu().numpy()
        assert np.array_equal(out[0], frames[0]) and np.array_equal(out[2], frames[2])
    except Exception as e:
        assert "packet 1" in str(e)
    with pytest.raises(Exception, match="packet 0"):
        dec.decode([packets[0][:-3]], rgb=rgb)


def test_reader_reads_writer_files_and_refuses_opencv_files(tmp_path):
    from metric_depth_video_toolbox_b200 import _lib, video_io

    w, h, n = 640, 360, 11
    frames = content(w, h, n, seed=21)
    path = str(tmp_path / "gpu.mkv")
    wr = ffv1_gpu.GpuFfv1Writer(path, 30.0, (w, h), device=DEV, batch=4)
    wr.write(torch.from_numpy(frames).to(DEV))
    assert wr.close() == n
    rd = ffv1_gpu.GpuFfv1Reader(path, device=DEV, batch=4)
    assert len(rd) == n and (rd.width, rd.height) == (w, h) and abs(rd.fps - 30.0) < 1e-6
    chunks = list(rd)
    assert [int(c.shape[0]) for c in chunks] == [4, 4, 3]
    assert np.array_equal(torch.cat(chunks).cpu().numpy(), frames)
    part = ffv1_gpu.GpuFfv1Reader(path, device=DEV, batch=8, start=3, stop=9).read_all()
    assert np.array_equal(part.cpu().numpy(), frames[3:9])
    cvpath = str(tmp_path / "cv.mkv")
    video_io.write_clip(cvpath, frames[:3], 30.0)
    with pytest.raises(_lib.MdvtError):
        ffv1_gpu.GpuFfv1Reader(cvpath, device=DEV)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_cli_stereo_rerender_gpu_writer_torchrun_two_ranks(tmp_path):
    """--gpu_ffv1 under torchrun: every rank codes its GOP-aligned frame range into a segment, rank 0 joins the segments at
    packet level; the joined files hold the frames of the single-process run."""
    import subprocess
    import sys

    import stereo_rerender
    from metric_depth_video_toolbox_b200 import video_io
    from metric_depth_video_toolbox_b200.synth import SyntheticClip

    w, h, n = 192, 108, 29            # 29 frames: ranges [0, 12) and [12, 29) with GOP-aligned starts
    depth, colour = SyntheticClip(w, h, n).frames()
    dpath, cpath = str(tmp_path / "depth.mkv"), str(tmp_path / "colour.mkv")
    video_io.write_clip(dpath, depth, 24.0)
    video_io.write_clip(cpath, colour, 24.0)
    base = ["--depth_video", dpath, "--color_video", cpath, "--xfov", "60", "--infill_mask", "--green_and_black_infill_mask",
            "--dont_place_points_in_edges", "--gpu_ffv1"]
    names = [dpath + "_stereo.mkv", dpath + "_stereo.mkv_infillmask.mkv"]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
           "29531", os.path.join(ROOT, "stereo_rerender.py")] + base
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    sharded = [video_io.read_clip(p) for p in names]
    assert not [f for f in os.listdir(tmp_path) if ".rank" in f or f.endswith(".plan.json")]
    for p in names:
        os.remove(p)
    assert stereo_rerender.main(base) == 0
    for p, got in zip(names, sharded):
        want = video_io.read_clip(p)
        assert got.shape == want.shape and got.shape[0] == n and np.array_equal(got, want), p


# ---- context model 1 (63 contexts): same kernels instantiated with the 5-level quantiser ------------------------------------
@pytest.mark.parametrize("w,h,slices,alpha,rgb", [(256, 144, (16, 9), False, True), (200, 120, (7, 5), True, False)])
def test_small_context_model_packets_equal_host_stepped_coder(host_coder, w, h, slices, alpha, rgb):
    test_packets_equal_host_stepped_coder(host_coder, w, h, slices, alpha, rgb, 1)


def test_small_context_model_files_decode_bit_exactly_in_opencv(tmp_path):
    test_writer_files_decode_bit_exactly_in_opencv(tmp_path, 640, 360, 9, 1)


@pytest.mark.parametrize("w,h,slices,alpha,rgb", [(256, 144, (16, 9), False, True), (200, 120, (7, 5), True, False)])
def test_small_context_model_decoder_mirrors_encoder(w, h, slices, alpha, rgb):
    test_device_decoder_mirrors_encoder(w, h, slices, alpha, rgb, 1)


# ---- context model 2 (14 contexts): the encoder keeps the coder states in shared memory (ffv1_encode_tiny_kernel; with the alpha
# plane the global-state kernel codes it) -------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h,slices,alpha,rgb", [(256, 144, (16, 9), False, True), (200, 120, (7, 5), True, False), (3840, 1080, None, False, True)])
def test_tiny_context_model_packets_equal_host_stepped_coder(host_coder, w, h, slices, alpha, rgb):
    if slices is None:
        slices = ffv1_gpu.slice_grid(w, h)
    test_packets_equal_host_stepped_coder(host_coder, w, h, slices, alpha, rgb, 2)


def test_tiny_context_model_files_decode_bit_exactly_in_opencv(tmp_path):
    test_writer_files_decode_bit_exactly_in_opencv(tmp_path, 640, 360, 9, 2)


@pytest.mark.parametrize("w,h,slices,alpha,rgb", [(256, 144, (16, 9), False, True), (200, 120, (7, 5), True, False)])
def test_tiny_context_model_decoder_mirrors_encoder(w, h, slices, alpha, rgb):
    test_device_decoder_mirrors_encoder(w, h, slices, alpha, rgb, 2)


def test_cli_stereo_rerender_with_gpu_writer_device_hand_off(tmp_path):
    """Without --create_sbs_depth_video the plain stereo mode hands the row kernel's device tensors to the coder."""
    test_cli_stereo_rerender_with_gpu_writer_equals_default_writer(tmp_path, depth_video=False)


def test_view_depthfile_and_save_depth_video_with_gpu_writer_equal_default_writers(tmp_path, monkeypatch):
    """3d_view_depthfile --render --gpu_ffv1 and depth_frames_helper.save_depth_video under MDVT_FFV1_WRITER=gpu write the
    same frames as their default (cv2.VideoWriter lanes) writers."""
    import importlib

    import depth_frames_helper as dfh
    from metric_depth_video_toolbox_b200 import video_io
    from metric_depth_video_toolbox_b200.synth import SyntheticClip

    w, h, n = 192, 108, 5
    depth, colour = SyntheticClip(w, h, n).frames()
    dpath, cpath = str(tmp_path / "depth.mkv"), str(tmp_path / "colour.mkv")
    video_io.write_clip(dpath, depth, 24.0)
    video_io.write_clip(cpath, colour, 24.0)
    view = importlib.import_module("3d_view_depthfile")
    argv = ["--depth_video", dpath, "--color_video", cpath, "--xfov", "60", "--render", "--chunk_frames", "3"]
    monkeypatch.setenv("MDVT_FFV1_WRITER", "host")
    assert view.main(argv) == 0
    monkeypatch.delenv("MDVT_FFV1_WRITER")
    want = video_io.read_clip(dpath + "_render.mkv")
    os.remove(dpath + "_render.mkv")
    assert view.main(argv + ["--gpu_ffv1"]) == 0
    got = video_io.read_clip(dpath + "_render.mkv")
    assert want.shape == (n, h, w, 3) and np.array_equal(got, want)

    rng = np.random.default_rng(5)
    metres = rng.uniform(0.2, 60, (n, 48, 64)).astype(np.float32)
    a, b = str(tmp_path / "host.mkv"), str(tmp_path / "gpu.mkv")
    monkeypatch.setenv("MDVT_FFV1_WRITER", "host")
    dfh.save_depth_video(metres, a, 24.0, 100, 64, 48)
    monkeypatch.delenv("MDVT_FFV1_WRITER")
    dfh.save_depth_video(metres, b, 24.0, 100, 64, 48)   # default with a device: coded there
    fa, fb = video_io.read_clip(a), video_io.read_clip(b)
    assert fa.shape == (n, 48, 64, 3) and np.array_equal(fa, fb)
    assert mkv_packets_all_key(b) and not mkv_packets_all_key(a)
    assert dfh.verify_and_move(b, n, str(tmp_path / "final.mkv"))


def mkv_packets_all_key(path):
    from metric_depth_video_toolbox_b200 import mkv_join

    pk = mkv_join.MkvPackets(path)
    try:
        return all(key for _, _, key in pk.packets)
    finally:
        pk.close()


def test_device_chunk_reader_yields_the_written_frames_and_falls_back_for_opencv_files(tmp_path):
    """video_io.open_chunk_reader: inputs written by this package's writer are decoded on the device (CUDA chunks, never a
    host pixel), in lock step over several inputs, with start / stop / chunking and the device BGR2GRAY; files written by
    cv2.VideoWriter (what the reference's tools produce) go through the host reader with the same frames."""
    from metric_depth_video_toolbox_b200 import video_io

    w, h, n = 320, 192, 11
    a, b = content(w, h, n, 1), content(w, h, n, 2)
    pa, pb, pc = (str(tmp_path / f"{k}.mkv") for k in "abc")
    for path, frames in ((pa, a), (pb, b)):
        wr = ffv1_gpu.GpuFfv1Writer(path, 24.0, (w, h), device=DEV, batch=4)
        wr.write(torch.from_numpy(frames).to(DEV))
        assert wr.close() == n
    rd = video_io.open_chunk_reader([pa, pb, None], 2, 10, chunk=3, grey=[False, True, False])
    assert isinstance(rd, video_io.DeviceChunkReader)
    got_a, got_b = [], []
    for cnt, (ca, cb, none) in rd:
        assert none is None and ca.is_cuda and cb.is_cuda and ca.shape[0] == cnt == cb.shape[0] and cb.dim() == 3
        got_a.append(ca.cpu().numpy())
        got_b.append(cb.cpu().numpy())
    assert np.array_equal(np.concatenate(got_a), a[2:10])
    want_grey = np.stack([cv2.cvtColor(cv2.cvtColor(f, cv2.COLOR_RGB2BGR), cv2.COLOR_BGR2GRAY) for f in b[2:10]])
    assert np.array_equal(np.concatenate(got_b), want_grey)
    # early exit: the background thread is stopped and joined by the generator's finally
    rd = video_io.open_chunk_reader([pa], chunk=2)
    for cnt, (ca,) in rd:
        break
    assert not rd._thread.is_alive()
    # a file from cv2.VideoWriter: host reader, same frames
    cvw = cv2.VideoWriter(pc, cv2.VideoWriter_fourcc(*"FFV1"), 24.0, (w, h))
    for f in a:
        cvw.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    cvw.release()
    rd = video_io.open_chunk_reader([pc, pa], chunk=4)   # mixed inputs: everything goes through the host reader
    assert isinstance(rd, video_io.ChunkReader)
    host = np.concatenate([c0.numpy().copy() for _, (c0, c1) in rd])
    assert np.array_equal(host, a)


def test_movie_steps_4_and_5_files_in_files_out_on_the_device(tmp_path):
    """movie_2_3D steps 4 + 5 on clips this package wrote: device decode -> render -> device encode (packets are the only
    pixels-related bytes on the host), and the results decode in OpenCV to exactly the frames of the host-I/O run."""
    import json

    from metric_depth_video_toolbox_b200 import movie_steps, video_io
    from metric_depth_video_toolbox_b200.cli import stereo_rerender
    from metric_depth_video_toolbox_b200.synth import SyntheticClip

    w, h, n = 320, 192, 14
    depth, colour = SyntheticClip(w, h, n, zero_fraction=0.003).frames()
    mask = np.zeros((n, h, w, 3), np.uint8)
    mask[:, 40:150, 60:260] = 255
    results = {}
    for mode in ("device", "host"):
        d = tmp_path / mode
        d.mkdir()
        paths = {k: str(d / f"{k}.mkv") for k in ("depth", "colour", "mask")}
        for key, frames in (("depth", depth), ("colour", colour), ("mask", mask)):
            if mode == "device":
                wr = ffv1_gpu.GpuFfv1Writer(paths[key], 24.0, (w, h), device=DEV, batch=4)
            else:
                wr = video_io.ChunkWriter(paths[key], "FFV1", 24.0, (w, h))
            wr.write(torch.from_numpy(frames), rgb=True)
            wr.close()
        os.environ["MDVT_FFV1_WRITER"] = "gpu" if mode == "device" else "host"
        try:
            scene = {"finished": False, "scene_video_file": paths["colour"], "depth_video_file": paths["depth"], "mask_video_file": paths["mask"],
                     "xfov": 60.0, "sbs": str(d / "sbs.mkv")}
            movie_steps.step4_find_convergence([scene])
            argv = movie_steps.stereo_rerender_argv(scene) + ["--green_and_black_infill_mask", "--create_sbs_depth_video"]
            stereo_rerender.run(stereo_rerender.build_parser().parse_args(argv), keep_process_group=True)
        finally:
            os.environ.pop("MDVT_FFV1_WRITER", None)
        conv = json.load(open(paths["depth"] + "_convergence_depths.json"))
        out = paths["depth"] + "_stereo.mkv"
        frames = {}
        for suffix in ("", "_infillmask.mkv", "_depth.mkv"):
            cap = cv2.VideoCapture(out + suffix)
            got = []
            while True:
                ok, f = cap.read()
                if not ok:
                    break
                got.append(f)
            cap.release()
            frames[suffix] = np.stack(got)
        results[mode] = (conv, frames)
    assert results["device"][0] == results["host"][0] and len(results["host"][0]) == n
    for suffix in ("", "_infillmask.mkv", "_depth.mkv"):
        assert results["device"][1][suffix].shape[0] == n
        assert np.array_equal(results["device"][1][suffix], results["host"][1][suffix]), suffix


@pytest.mark.parametrize("w,h,nh,nv,alpha,model", [(64, 48, 8, 8, False, 0), (70, 33, 5, 7, True, 0), (64, 48, 16, 12, False, 1), (48, 32, 1, 1, False, 1)])
def test_device_decoder_decodes_oracle_written_streams(w, h, nh, nv, alpha, model):
    """an anchor for the device DECODER that does not involve the device (or host-stepped) encoder: the
    packets come from oracle/ffv1_oracle.py's plain-Python encoder (pinned byte for byte on libavcodec), many slices, both
    context models, with and without the alpha plane, both channel orders."""
    from oracle import ffv1_oracle as fo

    base = fo.parse_config(ffv1_gpu.stream_setup(w, h, nh, nv, alpha, model)[0])
    frames = content(w, h, 4, seed=21)          # RGB-order test content; the oracle codes B, G, R, A planes
    packets = []
    for f in frames:
        bgra = np.dstack([f[..., ::-1], np.full((h, w), 255, np.uint8)])
        packets.append(fo.encode_frame(bgra, base, True, [fo.SliceState(base) for _ in range(nh * nv)]))
    dec = ffv1_gpu.Ffv1Decoder(w, h, DEV, max_frames=4, slices=(nh, nv), alpha=alpha, context_model=model)
    assert np.array_equal(dec.decode(packets, rgb=True).cpu().numpy(), frames)
    assert np.array_equal(dec.decode(packets[:2], rgb=False).cpu().numpy(), frames[:2, ..., ::-1])
