"""CPU: host-side logic of the drop-in surface -- command-line parity with the reference scripts (flag
metadata extracted from the reference into tests/golden/cli_flags.json), frame sharding and the parameter
broadcast over a world_size-2 gloo group, and the threaded clip reader / writer."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metric_depth_video_toolbox_b200 import sharding
from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer

TYPES = {"int": int, "float": float, "str": str, None: None}


def _parsers():
    from metric_depth_video_toolbox_b200.cli import convert_format, find_convergence_depth, stereo_rerender, view_depthfile

    return {"stereo_rerender.py": stereo_rerender.build_parser(), "3d_view_depthfile.py": view_depthfile.build_parser(),
            "convert_metric_depth_video_to_other_format.py": convert_format.build_parser(),
            "find_convergence_depth.py": find_convergence_depth.build_parser()}


def test_cli_flags_match_reference(golden_dir):
    golden = json.load(open(os.path.join(golden_dir, "cli_flags.json")))
    parsers = _parsers()
    assert set(parsers) == set(golden)
    for script, flags in golden.items():
        actions = {a.option_strings[0]: a for a in parsers[script]._actions if a.option_strings}
        for flag, spec in flags.items():
            assert flag in actions, f"{script}: {flag} missing"
            a = actions[flag]
            if spec["action"] == "store_true":
                assert a.nargs == 0 and a.const is True and a.default is False, (script, flag)
            else:
                assert a.type is TYPES[spec["type"]], (script, flag, a.type)
                assert a.default == spec["default"] and type(a.default) is type(spec["default"]), (script, flag, a.default)
            assert bool(a.required) == bool(spec["required"]), (script, flag)


def test_root_launchers_exist_with_reference_names():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name in ("stereo_rerender.py", "3d_view_depthfile.py", "convert_metric_depth_video_to_other_format.py", "find_convergence_depth.py",
                 "depth_frames_helper.py", "depth_map_tools.py"):
        assert os.path.isfile(os.path.join(root, name)), name


def test_stereo_rerender_validation_errors(tmp_path):
    from metric_depth_video_toolbox_b200.cli import stereo_rerender as sr

    with pytest.raises(ValueError, match="Either --xfov_file, --xfov or --yfov"):
        sr.main(["--depth_video", "nope.mkv"])
    with pytest.raises(ValueError, match="not compatible"):
        sr.main(["--depth_video", "nope.mkv", "--xfov", "60", "--green_and_black_infill_mask", "--do_basic_infill"])
    with pytest.raises(FileNotFoundError, match="Depth video not found"):
        sr.main(["--depth_video", str(tmp_path / "nope.mkv"), "--xfov", "60"])
    args = sr.build_parser().parse_args(["--depth_video", "a.mkv", "--xfov", "60"])
    assert sr.output_names(args) == ("a.mkv_stereo.mkv", "a.mkv_tmp_stereo.mkv", "FFV1")
    args = sr.build_parser().parse_args(["--depth_video", "a.mkv", "--xfov", "60", "--touchly1", "--compressed"])
    assert sr.output_names(args) == ("a.mkv_Touchly1.mp4", "a.mkv_tmp_Touchly1.mp4", "avc1")


# ---------------------------------------------------------------------------------------------
# sharding
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,world", [(0, 1), (1, 8), (7, 8), (8, 8), (300, 1), (2400, 8), (2401, 4), (1000, 3)])
def test_frame_ranges_partition_the_clip(n, world):
    ranges = [sharding.frame_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    with pytest.raises(ValueError):
        sharding.frame_range(n, world, world)


def _example_params(n):
    rng = np.random.default_rng(3)
    T = np.tile(np.eye(4), (n, 1, 1))
    T[:, :3, 3] = rng.normal(size=(n, 3))
    return StereoParams(1920, 1080, xfov=None, yfov=None, xfovs=list(rng.uniform(40, 80, n)), max_depth=100, pupillary_distance=63,
                        master_xfov=45.0, convergence_depths=rng.uniform(1, 9, n), transformations=T, infill_mask=True, mask_rgb=True)


def test_param_block_round_trip():
    p = _example_params(11)
    q, n = sharding.unpack_params(sharding.pack_params(p, 11))
    assert n == 11 and q.width == 1920 and q.height == 1080 and q.xfov is None and q.yfov is None and q.mask_rgb and q.infill_mask
    assert np.array_equal(q.xfovs, p.xfovs) and np.array_equal(q.convergence_depths, p.convergence_depths)
    assert np.array_equal(q.transformations, p.transformations)
    plain = StereoParams(640, 480, xfov=60.0, infill_mask=False)
    q, n = sharding.unpack_params(sharding.pack_params(plain, 5))
    assert (q.xfov, q.yfov, q.xfovs, q.convergence_depths, q.transformations, q.infill_mask) == (60.0, None, None, None, None, False)
    with pytest.raises(ValueError, match="entries"):
        sharding.pack_params(p, 12)
    sub = sharding.shard_params(p, 3, 7)
    assert len(sub.xfovs) == 4 and np.array_equal(sub.transformations, p.transformations[3:7])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, n_frames, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.world() == (rank, world)
        src = _example_params(n_frames) if rank == 0 else None
        params, n = sharding.broadcast_params(src, n_frames if rank == 0 else 0)
        start, stop = sharding.frame_range(n)
        # each rank derives the per-frame constant block of its own shard (host-side part of the render)
        consts = StereoRerenderer.__new__(StereoRerenderer)
        consts.p = params
        block = StereoRerenderer.frame_constants(consts, start, stop - start)
        views = StereoRerenderer.views_of(consts, start)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), start=start, stop=stop, block=block, M=views[0].M,
                 conv=np.asarray(params.convergence_depths), n=n)
        assert sharding.gather_counts(stop - start) == n
    finally:
        dist.destroy_process_group()


def test_gloo_world2_broadcast_and_shard(tmp_path):
    n = 9
    mp.spawn(_gloo_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(2))
    assert (int(r0["start"]), int(r0["stop"]), int(r1["start"]), int(r1["stop"])) == (0, 5, 5, 9)
    assert int(r1["n"]) == n
    whole = _example_params(n)
    rr = StereoRerenderer.__new__(StereoRerenderer)
    rr.p = whole
    want = StereoRerenderer.frame_constants(rr, 0, n)
    assert np.array_equal(np.concatenate([r0["block"], r1["block"]]), want)  # shards concatenate to the single-rank block
    assert np.array_equal(r1["conv"], whole.convergence_depths)             # rank 1 received the whole smoothed list
    assert np.array_equal(r1["M"], StereoRerenderer.views_of(rr, 5)[0].M)


# ---------------------------------------------------------------------------------------------
# clip reader / writer threads
# ---------------------------------------------------------------------------------------------
def test_chunk_reader_writer_round_trip(tmp_path):
    from metric_depth_video_toolbox_b200 import video_io

    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (7, 32, 48, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (7, 32, 48, 3), dtype=np.uint8)
    pa, pb = str(tmp_path / "a.mkv"), str(tmp_path / "b.mkv")
    video_io.write_clip(pa, a, 24.0)
    video_io.write_clip(pb, b, 24.0)
    assert video_io.video_info(pa) == (48, 32, 24.0, 7)
    assert np.array_equal(video_io.read_clip(pa), a)  # FFV1 is lossless
    got_a, got_b = [], []
    for n, (ca, cb, none) in video_io.ChunkReader([pa, pb, None], start=2, stop=7, chunk=2, pin=False):
        assert none is None and ca.shape == (n, 32, 48, 3)
        got_a.append(ca.numpy().copy())
        got_b.append(cb.numpy().copy())
    assert [len(x) for x in got_a] == [2, 2, 1]
    assert np.array_equal(np.concatenate(got_a), a[2:]) and np.array_equal(np.concatenate(got_b), b[2:])
    grey = [c[0].numpy().copy() for _, c in video_io.ChunkReader([pa], chunk=4, pin=False, grey=[True])]
    assert np.concatenate(grey).shape == (7, 32, 48)


def test_movie_step5_command_line_matches_reference_string():
    """movie_2_3D.py:431-445 builds `stereo_rerender.py --color_video S {conv} {xfov} --depth_video D {edge} {infm}`."""
    from metric_depth_video_toolbox_b200.cli import stereo_rerender
    from metric_depth_video_toolbox_b200.movie_steps import stereo_rerender_argv

    scene = {"scene_video_file": "s.mkv", "depth_video_file": "d.mkv", "convergence_file": "c.json", "xfov": 52.5}
    assert stereo_rerender_argv(scene) == ["--color_video", "s.mkv", "--convergence_file", "c.json", "--xfov", "52.5", "--depth_video", "d.mkv",
                                           "--infill_mask"]
    scene = {"scene_video_file": "s.mkv", "depth_video_file": "d.mkv", "xfovs_file": "x.json", "xfov": None, "infill": False, "convergence": False}
    argv = stereo_rerender_argv(scene)
    assert argv == ["--color_video", "s.mkv", "--xfov_file", "x.json", "--depth_video", "d.mkv"]
    args = stereo_rerender.build_parser().parse_args(argv)
    assert args.xfov_file == "x.json" and not args.infill_mask and args.convergence_file is None


def test_equirect_maps_match_reference(golden_dir):
    """The VR180 coordinate maps are host-side NumPy (they depend on the frame size and FOV only); with OpenCV's remap
    they must reproduce the reference's convert_to_equirectangular outputs bit for bit."""
    import cv2

    from metric_depth_video_toolbox_b200 import vr180

    g = np.load(os.path.join(golden_dir, "misc.npz"))
    for k in range(3):
        img, fov = g[f"equirect_in{k}"], float(g[f"equirect_fov{k}"])
        mx, my = vr180.equirect_maps(img.shape[0], img.shape[1], fov)
        assert mx.dtype == np.float32 and (mx == -1).any()
        out = cv2.remap(img, mx, my, interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=(0, 0, 0))
        assert np.array_equal(out, g[f"equirect_out{k}"]), k


@pytest.mark.parametrize("n,world,align", [(100, 3, 12), (7, 2, 12), (2400, 8, 12), (25, 4, 12), (12, 2, 12)])
def test_aligned_frame_ranges_partition_the_clip(n, world, align):
    ranges = [sharding.frame_range(n, r, world, align) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a0 <= a1
    assert all(a % align == 0 or a == n for a, _ in ranges)


def _noise_clip(n, h=48, w=64, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (n, h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("lanes,n,step", [(1, 30, 8), (3, 77, 8), (8, 25, 12), (4, 12, 5), (5, 3, 3)])
def test_parallel_writer_joins_lanes_into_the_same_frames(tmp_path, lanes, n, step):
    """K FFV1 encoder lanes + the packet-level Matroska join == the frames a single cv2.VideoWriter file holds:
    same frame count, fps and size as OpenCV reports them, every frame bit-identical, exact random access."""
    import cv2

    from metric_depth_video_toolbox_b200 import video_io

    frames = _noise_clip(n)
    path = str(tmp_path / "out.mkv")
    w = video_io.ParallelWriter(path, 24.0, (64, 48), lanes=lanes)
    for s in range(0, n, step):
        w.write(frames[s:s + step])
    w.close()
    assert sorted(os.listdir(tmp_path)) == ["out.mkv"]          # lane files are gone
    assert video_io.video_info(path) == (64, 48, 24.0, n)
    assert np.array_equal(video_io.read_clip(path), frames)
    cap = cv2.VideoCapture(path)
    for pos in (n - 1, n // 2, 0):
        cap.set(cv2.CAP_PROP_POS_FRAMES, pos)
        ok, f = cap.read()
        assert ok and np.array_equal(f[..., ::-1], frames[pos])


def test_parallel_writers_of_two_ranks_join_in_rank_order(tmp_path):
    from metric_depth_video_toolbox_b200 import video_io

    frames = _noise_clip(50, seed=1)
    parts = []
    for r, (a, b) in enumerate(((0, 24), (24, 50))):
        p = str(tmp_path / f"seg.rank{r:02d}.mkv")
        w = video_io.ParallelWriter(p, 30.0, (64, 48), lanes=3, join_on_close=False)
        w.write(frames[a:b])
        w.close()
        parts.append(p)
    out = str(tmp_path / "joined.mkv")
    assert video_io.join_plans([video_io.load_plan(p) for p in parts], out, 30.0) == 50
    assert sorted(os.listdir(tmp_path)) == ["joined.mkv"]
    assert video_io.video_info(out) == (64, 48, 30.0, 50)
    assert np.array_equal(video_io.read_clip(out), frames)


@pytest.mark.parametrize("decoders,chunk,start", [(1, 8, 0), (3, 12, 0), (3, 12, 24), (2, 24, 12), (3, 8, 5)])
def test_chunk_reader_parallel_decoders_read_the_same_frames(tmp_path, decoders, chunk, start):
    from metric_depth_video_toolbox_b200 import video_io

    a, b = _noise_clip(77, seed=2), _noise_clip(70, seed=3)
    pa, pb = str(tmp_path / "a.mkv"), str(tmp_path / "b.mkv")
    video_io.write_clip(pa, a, 24.0)
    w = video_io.ParallelWriter(pb, 24.0, (64, 48), lanes=3)   # a joined file as input
    w.write(b)
    w.close()
    r = video_io.ChunkReader([pa, pb], start, None, chunk=chunk, decoders=decoders, pin=False)
    assert r.decoders == decoders      # Matroska inputs: every decoder is handed the packets of its chunks, any chunk size / start
    got_a, got_b = [], []
    for n, (xa, xb) in r:
        got_a.append(xa.numpy().copy())
        got_b.append(xb.numpy().copy())
    assert np.array_equal(np.concatenate(got_a), a[start:70]) and np.array_equal(np.concatenate(got_b), b[start:])


def test_mkv_join_refuses_a_run_that_does_not_start_on_a_key_frame(tmp_path):
    from metric_depth_video_toolbox_b200 import mkv_join, video_io

    p = str(tmp_path / "a.mkv")
    video_io.write_clip(p, _noise_clip(20), 24.0)
    with pytest.raises(mkv_join.MkvError):
        mkv_join.join([(p, 5), (p, 5)], str(tmp_path / "bad.mkv"), 24.0)   # packet 5 is inside a GOP
    assert mkv_join.join([(p, 12), (p, 8)], str(tmp_path / "ok.mkv"), 24.0) == 20


@pytest.mark.parametrize("fps", [23.976, 29.97, 59.94, 12.5])
def test_parallel_writer_keeps_fractional_frame_rates(tmp_path, fps):
    """The joined file reports the frame rate and frame count of a single-writer file (verify_and_move checks the count)."""
    from metric_depth_video_toolbox_b200 import video_io

    frames = _noise_clip(40, seed=4)
    single, joined = str(tmp_path / "s.mkv"), str(tmp_path / "p.mkv")
    video_io.write_clip(single, frames, fps)
    w = video_io.ParallelWriter(joined, fps, (64, 48), lanes=3)
    w.write(frames)
    w.close()
    assert video_io.video_info(joined) == video_io.video_info(single) == (64, 48, fps, 40)
    assert np.array_equal(video_io.read_clip(joined), frames)


def _random_mask_image(h, w, seed, holes, max_width):
    """A pre-inpaint mask image like the GPU produces: black, green hole stripes, coded normals scattered in them."""
    rng = np.random.default_rng(seed)
    m = np.zeros((h, w, 3), np.uint8)
    for _ in range(holes):
        x, y = int(rng.integers(5, w - max_width - 5)), int(rng.integers(5, h - 40))
        hh, ww = int(rng.integers(8, min(120, h - y - 2))), int(rng.integers(2, max_width))
        m[y:y + hh, x:x + ww] = (0, 255, 0)
        pts = int(rng.integers(hh // 2, hh * 2))
        m[rng.integers(y, y + hh, pts), rng.integers(max(0, x - 1), min(w, x + ww + 1), pts)] = rng.integers(1, 255, (pts, 3)).astype(np.uint8)
    return m


@pytest.mark.parametrize("seed", range(8))
def test_restricted_telea_area_gives_the_reference_hole_pixels(seed, monkeypatch):
    """finish_mask inpaints only as far around the holes as can influence a hole pixel; the result must be the image the
    reference's full inpaint area (everything that is not a coded normal) gives, byte for byte."""
    from metric_depth_video_toolbox_b200 import infill

    m = _random_mask_image(150, 260, 50 + seed, holes=4 + 3 * seed, max_width=6 + 4 * seed)
    monkeypatch.delenv("MDVT_TELEA_FULL", raising=False)
    fast = infill.finish_mask(m)
    monkeypatch.setenv("MDVT_TELEA_FULL", "1")
    full = infill.finish_mask(m)
    assert np.array_equal(fast, full)
    assert (fast != m).any()


def test_restricted_telea_matches_the_reference_golden_masks():
    """The golden pre-inpaint / final mask pairs were produced by the reference's own lines (stereo_rerender.py:740-819)."""
    from metric_depth_video_toolbox_b200 import infill

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "infill_mask.npz"))
    for tag in ("plain", "posed"):
        assert np.array_equal(infill.finish_mask(g[f"{tag}_mask_pre_inpaint"]), g[f"{tag}_mask_final"]), tag
    blank = np.zeros((40, 60, 3), np.uint8)
    blank[5, 7] = (200, 100, 50)
    assert np.array_equal(infill.finish_mask(blank), infill.masked_blur(blank))   # no hole: nothing to inpaint


def test_infill_finish_async_equals_finish_per_eye():
    """The deferred (worker-pool) TELEA + blur tail produces, per eye, exactly what finish_mask gives."""
    from metric_depth_video_toolbox_b200 import infill

    frames = np.stack([np.concatenate([_random_mask_image(90, 128, 10 + k, 5, 12), _random_mask_image(90, 128, 20 + k, 7, 9)], axis=1)
                       for k in range(3)])
    im = infill.InfillMaskRenderer(None, workers=4)
    deferred = im.finish_async(frames)
    frames_before = frames.copy()
    frames[:] = 0                      # the caller recycles its buffer right away
    out = deferred.result()
    for k in range(3):
        for e in range(2):
            assert np.array_equal(out[k, :, e * 128:(e + 1) * 128], infill.finish_mask(frames_before[k, :, e * 128:(e + 1) * 128]))


def test_conv_vrows_limits_check_runs_on_the_host():
    """mdvt_stereo_conv_vrows_supported makes no CUDA call: typical convergence distances pass at the sizes the kernel is built
    for, a width that is not a multiple of 32, a pose that is not `y rotation + x shift`, and a convergence distance of a few
    centimetres (staircase steeper than 0.4 rows per 15 columns) do not."""
    import numpy as np

    from metric_depth_video_toolbox_b200 import ops
    from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer

    def frames(w, h, convs):
        rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=convs), "cpu")
        return ops.conv_frames_packed(*rr.packed_cameras(0, len(convs)), rr.p.near)

    assert ops.conv_vrows_supported(frames(1920, 1080, [5.0, 1.1, 0.75, 0.0]), 1920, 1080)
    assert ops.conv_vrows_supported(frames(3840, 2160, [4.0, 1.5]), 3840, 2160)
    assert not ops.conv_vrows_supported(frames(640, 480, [0.05]), 640, 480)
    assert not ops.conv_vrows_supported(frames(70, 33, [3.0]), 70, 33)
    assert not ops.conv_vrows_supported(frames(4096, 64, [3.0]), 4096, 64)
    tilted = frames(640, 480, [3.0])
    tilted[0, 8 + 9] = 1e-3  # M[9] of the left eye: a rotation about x sneaks in (the target row would depend on the depth)
    assert not ops.conv_vrows_supported(tilted, 640, 480)
    with pytest.raises(ValueError):
        ops.conv_vrows_supported(np.zeros((2, 7), dtype=np.float32), 640, 480)


def test_writer_context_model_comes_from_the_environment(monkeypatch):
    """MDVT_FFV1_CONTEXT_MODEL picks the device writers' quant tables (default 1; 2 = the 14-context tables); anything else is refused."""
    from metric_depth_video_toolbox_b200 import ffv1_gpu

    monkeypatch.delenv("MDVT_FFV1_CONTEXT_MODEL", raising=False)
    assert ffv1_gpu.default_context_model() == 1
    for v in ("0", "1", "2"):
        monkeypatch.setenv("MDVT_FFV1_CONTEXT_MODEL", v)
        assert ffv1_gpu.default_context_model() == int(v)
    monkeypatch.setenv("MDVT_FFV1_CONTEXT_MODEL", "3")
    with pytest.raises(ValueError):
        ffv1_gpu.default_context_model()
