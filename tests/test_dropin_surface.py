"""The drop-in surface the reference's OTHER scripts import ("signatures stay importable"): every
`def` of the reference's depth_map_tools (names extracted from the reference with `ast` into the list below by
oracle/make_dropin_golden.py's author; 37 of them), the module-level helpers of stereo_rerender that
basic_nomal_infill.py:10 / stereo_dissoclusion_net_infill.py:10 import, and the small host-side matrix helpers against
values produced by RUNNING the reference (tests/golden/dropin_helpers.npz).  The GPU half checks the per-pixel ones
(calculate_normals, infill_using_normals, create_mesh_from_point_cloud) bit for bit."""
import inspect
import os

import numpy as np
import pytest
import torch

import depth_map_tools as dmt
import stereo_rerender as sr

# every top-level `def` of /root/reference/depth_map_tools.py (:14-1658), in file order
DEPTH_MAP_TOOLS_DEFS = [
    "timer", "calculate_normals", "open_cv_w2c_to_gl_view", "frustum_planes", "frusta_intersect", "apply_side_view_to_paralax_mask",
    "rotation_y", "translation_matrix", "get_cam_view", "convergence_angle", "mesh_from_depth_and_rgb",
    "mesh_maker_helper_make_corner_unclamped", "mesh_maker_helper_make_corner_with_mask", "remap_ids_to_img", "steep_disparity_lr",
    "steep_mask_disparity", "generate_normal_bg_image", "gl_render", "open_gl_projection_from_camera_matrix", "compute_camera_matrix", "svd",
    "transform_points", "pnpSolve_ransac", "reject_outliers", "pts_2_pcd", "project_3d_points_to_2d", "project_2d_points_to_3d",
    "convert_mesh_to_pcd", "get_mesh_from_depth_map", "create_point_cloud_from_depth", "perspective_aware_down_sample",
    "create_mesh_from_point_cloud", "render", "gl_look_at", "cam_look_at", "fov_from_camera_matrix", "draw"]
DEPTH_MAP_TOOLS_GLOBALS = ["vis", "v_h", "v_w", "rend", "use_ofscreen", "zero_identity_matrix"]
STEREO_RERENDER_DEFS = ["timer", "convert_to_equirectangular", "make_infill_mask", "convergence_angle", "masked_blur", "infill_using_normals",
                        "fill_nan_with_closest", "curve_fit"]
OUTSIDE_THE_PATH = ["mesh_from_depth_and_rgb", "mesh_maker_helper_make_corner_unclamped", "mesh_maker_helper_make_corner_with_mask",
                    "remap_ids_to_img", "steep_disparity_lr", "steep_mask_disparity", "generate_normal_bg_image", "gl_render",
                    "open_gl_projection_from_camera_matrix", "frustum_planes", "frusta_intersect", "svd", "pnpSolve_ransac",
                    "project_2d_points_to_3d", "perspective_aware_down_sample", "draw"]


def test_depth_map_tools_name_coverage_37_of_37():
    assert len(DEPTH_MAP_TOOLS_DEFS) == 37
    missing = [n for n in DEPTH_MAP_TOOLS_DEFS if not callable(getattr(dmt, n, None))]
    assert missing == []
    assert [n for n in DEPTH_MAP_TOOLS_GLOBALS if not hasattr(dmt, n)] == []


def test_reference_def_list_is_current():
    """Where the reference checkout is reachable (this container, not the GPU box) the list above is re-derived from it."""
    from oracle import ref_bridge

    if not ref_bridge.available():
        pytest.skip("no reference checkout")
    import ast

    tree = ast.parse(open(os.path.join(ref_bridge.REFERENCE_ROOT, "depth_map_tools.py")).read())
    assert [n.name for n in tree.body if isinstance(n, ast.FunctionDef)] == DEPTH_MAP_TOOLS_DEFS
    tree = ast.parse(open(os.path.join(ref_bridge.REFERENCE_ROOT, "stereo_rerender.py")).read())
    assert [n.name for n in tree.body if isinstance(n, ast.FunctionDef)] == STEREO_RERENDER_DEFS
    # argument names / defaults of the functions this round added
    ref = ref_bridge.load("depth_map_tools")
    def params(fn):  # (name, default) per parameter: annotations are not part of the calling convention
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]

    for name in ("calculate_normals", "get_cam_view", "create_mesh_from_point_cloud", "gl_look_at", "open_cv_w2c_to_gl_view", "reject_outliers",
                 "apply_side_view_to_paralax_mask"):
        assert params(getattr(dmt, name)) == params(getattr(ref, name)), name
    ref_sr = ref_bridge.load("stereo_rerender")
    for name in ("infill_using_normals", "make_infill_mask", "masked_blur"):
        assert params(getattr(sr, name)) == params(getattr(ref_sr, name)), name


def test_imports_of_the_reference_infill_scripts_resolve():
    from stereo_rerender import infill_using_normals, masked_blur  # basic_nomal_infill.py:10, stereo_dissoclusion_net_infill.py:10

    assert callable(infill_using_normals) and callable(masked_blur)
    assert [n for n in STEREO_RERENDER_DEFS if not callable(getattr(sr, n, None))] == []
    assert sr.make_infill_mask(np.zeros((2, 2), bool), None) is None  # the reference's placeholder returns None (:89-91)


def test_names_outside_the_path_refuse_with_a_reason():
    for name in OUTSIDE_THE_PATH:
        with pytest.raises(NotImplementedError, match="not part of the GPU per-frame path"):
            getattr(dmt, name)()


def test_host_matrix_helpers_match_reference_outputs(golden_dir, capsys):
    g = np.load(os.path.join(golden_dir, "dropin_helpers.npz"))
    for got, want in ((dmt.get_cam_view(0.0315, 0.0063), g["cam_view_fwd"]), (dmt.get_cam_view(-0.0315, 0.02, reverse=True), g["cam_view_rev"]),
                      (dmt.gl_look_at(np.array([1.0, 2.0, 3.0], dtype=np.float32), np.array([0.5, -1.0, -4.0], dtype=np.float32),
                                      np.array([0.0, 1.0, 0.0], dtype=np.float32)), g["gl_look_at"]),
                      (dmt.open_cv_w2c_to_gl_view(g["w2c_in"]), g["w2c_gl"])):
        assert got.dtype == want.dtype and np.array_equal(got, want)
    assert np.array_equal(dmt.reject_outliers(g["outlier_in"], 1.5), g["outlier_mask"])
    assert np.array_equal(dmt.apply_side_view_to_paralax_mask(g["side_pm"], g["side_n"], True), g["side_right"])
    assert np.array_equal(dmt.apply_side_view_to_paralax_mask(g["side_pm"], g["side_n"], False), g["side_left"])
    with dmt.timer("block"):
        pass
    assert capsys.readouterr().out.startswith("block: ")


# ---------------------------------------------------------------------------------------------
# GPU: the per-pixel ones
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_calculate_normals_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "dropin_helpers.npz"))
    for d, k, n in (("cn_depth", "cn_K", "cn_normals"), ("cn_depth2", "cn_K2", "cn_normals2")):
        got = dmt.calculate_normals(g[d], g[k])
        assert got.dtype == np.float32 and got.shape == g[n].shape
        assert np.array_equal(got.view(np.uint32), g[n].view(np.uint32))
    dev = dmt.calculate_normals(torch.from_numpy(g["cn_depth"]).cuda(), g["cn_K"])  # CUDA in -> CUDA out
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy().view(np.uint32), g["cn_normals"].view(np.uint32))


@pytest.mark.gpu
def test_infill_using_normals_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "dropin_helpers.npz"))
    img = g["iun_img"].copy()
    got = sr.infill_using_normals(img, g["iun_hole"], g["iun_normals"])
    assert np.array_equal(got, g["iun_out"]) and np.array_equal(img, g["iun_img"])   # result equal, input untouched
    assert (got != g["iun_img"]).any()
    assert np.array_equal(sr.infill_using_normals(img, g["iun_hole"], g["iun_normals"], max_steps=12), g["iun_out_12"])
    t = sr.infill_using_normals(torch.from_numpy(img).cuda(), torch.from_numpy(g["iun_hole"]).cuda(), torch.from_numpy(g["iun_normals"]).cuda())
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), g["iun_out"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["plain", "posed"])
def test_create_mesh_from_point_cloud_matches_reference_mesh_builder(golden_dir, tag):
    """create_mesh_from_point_cloud on the points create_point_cloud_from_depth returns == what the reference's own mesh
    builder returned for that frame (tests/golden/infill_mask.npz, made by oracle/make_infill_golden.py)."""
    from oracle import mdvt_oracle as orc

    g = np.load(os.path.join(golden_dir, "infill_mask.npz"))
    w, h, xfov, _ = g[tag + "_params"]
    w, h = int(w), int(h)
    K = orc.camera_matrix(xfov, None, w, h)
    depth = orc.apply_depth_scale(orc.decode_rgb_depth_frame(g[tag + "_depth_rgb"], 100, True), orc.master_fov_depth_scale(45.0, xfov))
    points, hh, ww = dmt.create_point_cloud_from_depth(depth, K, True)
    mesh, unused, removed = dmt.create_mesh_from_point_cloud(points, hh, ww, g[tag + "_colour"], None, True, return_normals_of_removed=True)
    assert np.array_equal(unused, g[tag + "_unused"]) and np.abs(removed - g[tag + "_removed_normals"]).max() < 1e-14
    assert np.array_equal(mesh.vertices, points) and mesh.vertex_colors.shape == (w * h, 3)
    mesh2, used = dmt.create_mesh_from_point_cloud(points, hh, ww, remove_edges=True)
    assert np.array_equal(np.setdiff1d(np.arange(w * h), used), g[tag + "_unused"])
    mesh3, used3 = dmt.create_mesh_from_point_cloud(points, hh, ww)
    assert np.array_equal(used3, np.arange(w * h))


@pytest.mark.gpu
def test_encode_depth_float64_input_matches_reference(golden_dir):
    """a float64 depth array must be clipped and scaled in float64 (no early float32 rounding)."""
    import depth_frames_helper as dfh

    g = np.load(os.path.join(golden_dir, "dropin_helpers.npz"))
    d64 = g["enc64_depth"]
    differs = 0
    for md in (100, 20):
        codes = dfh.encode_depth_as_uint32(d64, md)
        assert codes.dtype == np.uint32 and np.array_equal(codes, g[f"enc64_codes_md{md}"])
        assert np.array_equal(dfh.encode_data_as_BGR(codes, 48, 32, bit16=True), g[f"enc64_bgr16_md{md}"])
        differs += int((dfh.encode_depth_as_uint32(d64.astype(np.float32), md) != codes).sum())
    assert differs > 0   # the float32 route really gives other codes: the test distinguishes the two


# ---------------------------------------------------------------------------------------------
# movie_2_3D: the module surface MDVT_gui.py:1290-1320 loads by path
# ---------------------------------------------------------------------------------------------
# every top-level `def` of /root/reference/movie_2_3D.py, in file order
MOVIE_2_3D_DEFS = [
    "write_frames_to_file", "wait_for_first", "is_valid_video", "validate_video_lengths", "_seconds_to_timecode", "split_scenes", "parse_args",
    "ensure_output_dir", "ensure_scene_file", "open_input_video", "load_and_split_scenes", "plan_scene_files", "step1_create_scene_videos",
    "step2_estimate_depth", "step3_generate_masks", "step4_find_convergence", "step5_render_sbs", "step6_infill_and_collect",
    "step6_normal_infill_render_sbs", "step6_m2svid_infill_and_collect", "step6_inspatio_world_infill_and_collect",
    "step6_stereocrafter_infill_and_collect", "step6_stereo_dissoclusion_net_infill_and_collect", "step7_concat_and_mux", "main"]


def _movie_golden():
    import json

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "movie_2_3D.json")) as fh:
        return json.load(fh)


def test_movie_2_3D_name_coverage_and_signatures():
    import movie_2_3D as m

    assert len(MOVIE_2_3D_DEFS) == 25
    assert [n for n in MOVIE_2_3D_DEFS if not callable(getattr(m, n, None))] == []
    from oracle import ref_bridge

    if ref_bridge.available():   # this container only: names, parameter names and defaults re-derived from the reference
        import ast

        tree = ast.parse(open(os.path.join(ref_bridge.REFERENCE_ROOT, "movie_2_3D.py")).read())
        defs = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
        assert [n.name for n in defs] == MOVIE_2_3D_DEFS
        for node in defs:
            want = [a.arg for a in node.args.args]
            got = list(inspect.signature(getattr(m, node.name)).parameters)
            assert got == want, (node.name, got, want)
            want_defaults = [ast.literal_eval(d) for d in node.args.defaults]
            got_defaults = [p.default for p in inspect.signature(getattr(m, node.name)).parameters.values() if p.default is not inspect.Parameter.empty]
            assert got_defaults == want_defaults, node.name


def test_movie_2_3D_planning_helpers_equal_the_reference_outputs(tmp_path):
    """Time codes, scene splitting, the csv loader and the per-scene file plan against the outputs of the reference's own
    functions (tests/golden/movie_2_3D.json, oracle/make_movie_golden.py)."""
    import copy

    import movie_2_3D as m

    g = _movie_golden()
    for text, want in g["timecodes"].items():
        assert m._seconds_to_timecode(float(text)) == want, text
    for max_frames, want in g["split"].items():
        scenes = copy.deepcopy(g["scenes"])
        assert m.split_scenes(scenes, max_scene_frames=int(max_frames)) == want
        assert scenes == g["scenes"]   # the input rows are not modified
    assert m.split_scenes(copy.deepcopy(g["scenes"])) == g["split"]["1500"]   # default max_scene_frames
    csv_path = tmp_path / "scenes.csv"
    csv_path.write_text(g["csv_text"], newline="")
    for max_frames, want in g["csv"].items():
        assert m.load_and_split_scenes(str(csv_path), ",", int(max_frames)) == want
    out_dir = str(tmp_path / "out")
    for end_scene, want in g["plan"].items():
        planned = m.plan_scene_files(m.split_scenes(copy.deepcopy(g["scenes"]), 1500), out_dir, int(end_scene))
        assert [{k: (v.replace(out_dir, "<OUT>") if isinstance(v, str) else v) for k, v in s.items()} for s in planned] == want
    # 'finished' follows the files on disk
    m.ensure_output_dir(out_dir)
    m.ensure_output_dir(out_dir)
    scenes = m.split_scenes(copy.deepcopy(g["scenes"]), 1500)
    open(os.path.join(out_dir, "scene_2.mkv_depth.mkv_stereo.mkv"), "wb").close()
    assert [s["finished"] for s in m.plan_scene_files(scenes, out_dir, -1)] == [False, True, False, False, False, False]


def test_movie_2_3D_flag_table_equals_the_reference(monkeypatch):
    import ast

    import movie_2_3D as m
    from metric_depth_video_toolbox_b200 import movie_plan

    sys_path = os.path.abspath(movie_plan.__file__)
    got = {}
    for node in ast.walk(ast.parse(open(sys_path).read())):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument":
            spec = {"type": None, "default": None, "action": None, "required": False}
            for kw in node.keywords:
                if kw.arg == "type":
                    spec["type"] = kw.value.id
                elif kw.arg in ("default", "action", "required"):
                    try:
                        spec[kw.arg] = ast.literal_eval(kw.value)
                    except ValueError:
                        spec[kw.arg] = "expr:" + ast.unparse(kw.value)
            got[ast.literal_eval(node.args[0])] = spec
    assert got == _movie_golden()["flags"]
    monkeypatch.setattr("sys.argv", ["movie_2_3D.py"])
    with pytest.raises(ValueError, match="need --color_video"):
        m.parse_args()
    monkeypatch.setattr("sys.argv", ["movie_2_3D.py", "--color_video", "in.mp4", "--end_scene", "3", "--infill_engine", "none"])
    a = m.parse_args()
    assert (a.color_video, a.end_scene, a.infill_engine, a.output_dir, a.max_scene_frames, a.no_render, a.gui) == ("in.mp4", 3, "none", "output", 1500, False, False)
    assert a.parallel == int(os.cpu_count() // 2)


def test_movie_2_3D_step1_copies_the_scene_frames_and_the_model_steps_refuse(tmp_path):
    import cv2

    import movie_2_3D as m
    from metric_depth_video_toolbox_b200 import video_io
    from metric_depth_video_toolbox_b200.movie_plan import OutOfScope

    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, size=(9, 32, 48, 3), dtype=np.uint8)
    src = str(tmp_path / "movie.mkv")
    video_io.write_clip(src, frames, 24.0)
    assert m.is_valid_video(src) and not m.is_valid_video(str(tmp_path / "nothing.mkv"))
    small = tmp_path / "small.mkv"
    small.write_bytes(b"x" * 2047)
    assert not m.is_valid_video(str(small))
    scenes = [{"Scene Number": "1", "Length (frames)": "4", "finished": False, "scene_video_file": str(tmp_path / "scene_1.mkv")},
              {"Scene Number": "2", "Length (frames)": "2", "finished": True, "scene_video_file": str(tmp_path / "scene_2.mkv")},
              {"Scene Number": "3", "Length (frames)": "3", "finished": False, "scene_video_file": str(tmp_path / "scene_3.mkv")}]
    cap, w, h, fps = m.open_input_video(src)
    assert (w, h) == (48, 32) and abs(fps - 24.0) < 1e-6
    m.step1_create_scene_videos(cap, scenes, fps, w, h)
    cap.release()
    assert np.array_equal(video_io.read_clip(scenes[0]["scene_video_file"]), frames[0:4])
    assert not os.path.exists(scenes[1]["scene_video_file"])                      # finished scenes are only read past
    assert np.array_equal(video_io.read_clip(scenes[2]["scene_video_file"]), frames[6:9])
    assert not os.path.exists(scenes[0]["scene_video_file"] + "_tmp.mkv")
    for s in scenes:
        s["infilled"] = s["scene_video_file"]
    assert m.validate_video_lengths([scenes[0], scenes[2]]) and not m.validate_video_lengths(scenes)
    args = type("A", (), {"infill_engine": "stereocrafter", "parallel": 1})()
    for call in (lambda: m.step2_estimate_depth(args, scenes), lambda: m.step3_generate_masks(args, scenes),
                 lambda: m.step6_infill_and_collect(args, scenes), lambda: m.step6_normal_infill_render_sbs(args, scenes),
                 lambda: m.step7_concat_and_mux(args, []), m.main):
        with pytest.raises(OutOfScope, match="outside the dense per-frame path"):
            call()
    args.infill_engine = "magic"
    with pytest.raises(Exception, match="unknown infill engine: magic"):
        m.step6_infill_and_collect(args, scenes)
    import subprocess
    import sys

    procs = [subprocess.Popen([sys.executable, "-c", "import time; time.sleep(%s)" % t]) for t in ("5", "0")]
    left = m.wait_for_first(procs)
    assert left == [procs[0]] and m.wait_for_first([]) == []
    procs[0].kill()
    procs[0].wait()


# ---------------------------------------------------------------------------------------------
# convert_metric_depth_video_to_other_format: module-level helpers
# ---------------------------------------------------------------------------------------------
CONVERT_DEFS = ["float_image_to_byte_image", "compute_weights_chunked", "best_intersection_point_vectorized_weighted", "find_nearby_points",
                "merge_global_points", "add_open3d_mesh", "add_point_cloud", "assign_vertex_color_material", "create_camera_alembic",
                "estimate_scale_shift"]


def test_convert_script_helpers_importable_and_the_numpy_ones_equal_the_reference(golden_dir):
    import convert_metric_depth_video_to_other_format as conv
    from metric_depth_video_toolbox_b200.convert_helpers import OutOfScope

    assert [n for n in CONVERT_DEFS if not callable(getattr(conv, n, None))] == [] and callable(conv.main)
    g = np.load(os.path.join(golden_dir, "convert_helpers.npz"))
    assert np.array_equal(conv.float_image_to_byte_image(g["img"]), g["bytes_default"])
    assert np.array_equal(conv.float_image_to_byte_image(g["img"], max_value=5.0, scale=200, log_scale=2), g["bytes_custom"])
    np.testing.assert_allclose(conv.estimate_scale_shift(g["depth"], g["target"]), g["scale_shift"], rtol=1e-12)
    for call in (lambda: conv.compute_weights_chunked(np.zeros((2, 3))), lambda: conv.merge_global_points([], []),
                 lambda: conv.create_camera_alembic([], "x.abc"), lambda: conv.add_point_cloud(np.zeros((1, 3)))):
        with pytest.raises(OutOfScope, match="outside the dense per-frame path"):
            call()
    from oracle import ref_bridge

    if ref_bridge.available():   # this container only: names, parameter names and defaults re-derived from the reference
        import ast

        tree = ast.parse(open(os.path.join(ref_bridge.REFERENCE_ROOT, "convert_metric_depth_video_to_other_format.py")).read())
        defs = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
        assert [n.name for n in defs] == CONVERT_DEFS
        for node in defs:
            assert list(inspect.signature(getattr(conv, node.name)).parameters) == [a.arg for a in node.args.args], node.name
            got_defaults = [p.default for p in inspect.signature(getattr(conv, node.name)).parameters.values() if p.default is not inspect.Parameter.empty]
            assert got_defaults == [ast.literal_eval(d) for d in node.args.defaults], node.name
