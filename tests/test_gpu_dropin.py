"""GPU: the drop-in surface (depth_frames_helper / depth_map_tools modules, the script front ends) against
the golden vectors generated from the reference and against the CPU oracle.  Everything goes NumPy ->
drop-in function -> C ABI -> CUDA kernel -> NumPy, as a user of the reference modules would call it."""
import json
import os

import numpy as np
import pytest
import torch

import depth_frames_helper as dfh   # the root-level launchers: same import names as the reference
import depth_map_tools as dmt
from metric_depth_video_toolbox_b200 import ops, video_io
from metric_depth_video_toolbox_b200.novel_view import NovelViewParams, NovelViewRenderer
from metric_depth_video_toolbox_b200.synth import SyntheticClip
from oracle import kernel_model as km
from oracle import mdvt_oracle as orc
from test_kernel_model import assert_differs_only_where_explained, explained_map

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


# ---------------------------------------------------------------------------------------------
# depth_frames_helper
# ---------------------------------------------------------------------------------------------
def test_depth_frames_helper_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "decode_all_codes.npz"))
    rgb = g["rgb"]
    h, w = rgb.shape[:2]
    for md in (100, 20):
        got = dfh.decode_rgb_depth_frame(rgb, md, True)
        assert isinstance(got, np.ndarray) and got.dtype == np.float32
        assert np.array_equal(bits(got), bits(g[f"d1_depth_md{md}"]))
        assert np.array_equal(bits(dfh.decode_rgb_depth_frame(rgb, md, False)), bits(g[f"d1_depth24_md{md}"]))
    codes = dfh.decode_rgb_as_data(rgb, w, h, True)
    assert codes.dtype == np.uint32 and np.array_equal(codes, g["d1_codes"])
    assert np.array_equal(dfh.decode_rgb_as_data(rgb, w, h, False), g["d1_codes24"])
    assert np.array_equal(bits(dfh.decode_uint32_as_depth(codes, 100)), bits(g["d1_depth_md100"]))
    e = np.load(os.path.join(golden_dir, "encode.npz"))
    for md in (100, 20):
        enc = dfh.encode_depth_as_uint32(e["depth"], md)
        assert enc.dtype == np.uint32 and np.array_equal(enc, e[f"codes_md{md}"])
        assert np.array_equal(dfh.encode_data_as_BGR(enc, 64, 48, bit16=True), e[f"bgr16_md{md}"])
        assert np.array_equal(dfh.encode_data_as_BGR(enc, 64, 48, bit16=False), e[f"bgr24_md{md}"])
    # CUDA tensors in -> CUDA tensors out, same bits
    dev = dfh.decode_rgb_depth_frame(cu(rgb), 100, True)
    assert dev.is_cuda and np.array_equal(bits(dev.cpu().numpy()), bits(g["d1_depth_md100"]))


def test_save_depth_video_round_trip(tmp_path):
    rng = np.random.default_rng(5)
    frames = rng.uniform(0.2, 60, (5, 48, 64)).astype(np.float32)
    path = str(tmp_path / "depth.mkv")
    dfh.save_depth_video(frames, path, 24.0, 100, 64, 48)
    back = video_io.read_clip(path)  # RGB order, as the scripts see it
    assert back.shape == (5, 48, 64, 3)
    for k in range(5):
        assert np.array_equal(back[k], orc.encode_depth_frame_rgb(frames[k], 100, True))
        assert np.abs(dfh.decode_rgb_depth_frame(back[k], 100, True) - frames[k]).max() < 1.56e-3
    assert dfh.verify_and_move(path, 5, str(tmp_path / "final.mkv")) and os.path.isfile(tmp_path / "final.mkv")
    assert dfh.verify_and_move(str(tmp_path / "final.mkv"), 6, str(tmp_path / "x.mkv")) is False


# ---------------------------------------------------------------------------------------------
# depth_map_tools
# ---------------------------------------------------------------------------------------------
def _frame(w=64, h=48, k=0, zero_fraction=0.005):
    depth_rgb, colour = SyntheticClip(w, h, 4, zero_fraction=zero_fraction).frame(k)
    return depth_rgb, colour, orc.decode_rgb_depth_frame(depth_rgb, 100, True)


@pytest.mark.parametrize("of_by_one", [False, True])
def test_create_point_cloud_from_depth_bit_exact(of_by_one):
    _, _, depth = _frame(640, 480)
    K = dmt.compute_camera_matrix(60.0, None, 640, 480)
    pts, h, w = dmt.create_point_cloud_from_depth(depth, K, of_by_one)
    assert (h, w) == (480, 640) and pts.dtype == np.float64
    assert np.array_equal(pts.view(np.uint64), orc.unproject(depth, K, of_by_one).view(np.uint64))


def test_transform_and_project_points(golden_dir):
    g = np.load(os.path.join(golden_dir, "geometry_64x48.npz"))
    rng = np.random.default_rng(2)
    pts = rng.normal(size=(1000, 3)) * [2, 2, 1] + [0, 0, 6]
    T = np.eye(4)
    T[:3, :3] = orc.rot_y(0.3)
    T[:3, 3] = (0.5, -0.25, 1.5)
    got = dmt.transform_points(pts, T)
    np.testing.assert_allclose(got, orc.apply_pose(pts, T), rtol=1e-13, atol=1e-14)
    K = dmt.compute_camera_matrix(60.0, 45.0, 640, 480)
    uv = dmt.project_3d_points_to_2d(got, K)
    K32 = K.astype(np.float32).astype(np.float64)
    u, v, _ = orc.project(got, K32)
    np.testing.assert_allclose(uv, np.stack((u, v), -1), rtol=1e-12, atol=1e-9)
    # the reference's own outputs for these two functions (cv2.projectPoints twin: <= 1e-4 px)
    np.testing.assert_allclose(dmt.transform_points(g["xyz_obo0"], g["T"]), g["xyz_T"], rtol=1e-12, atol=1e-12)
    ok = g["xyz_T"][:, 2] > 1e-4  # cv2 mirrors / special-cases z <= 0; those points are culled before projection here
    assert np.abs(dmt.project_3d_points_to_2d(g["xyz_T"], g["K"])[ok] - g["uv_T"][ok]).max() < 1e-4


def test_depth_mesh_pose_algebra_and_center():
    depth_rgb, colour, depth = _frame(64, 48)
    K = dmt.compute_camera_matrix(60.0, None, 64, 48)
    mesh, used = dmt.get_mesh_from_depth_map(depth, K, colour, None, of_by_one=True)
    assert len(used) == 64 * 48 and np.array_equal(used, np.arange(64 * 48))
    want = orc.unproject(depth, K, True)
    assert np.array_equal(mesh.vertices, want)
    np.testing.assert_allclose(mesh.get_center(), want.mean(axis=0), rtol=1e-11, atol=1e-12)
    assert np.array_equal(mesh.vertex_colors, colour.reshape(-1, 3) / 255.0)
    # the script's eye-pose sequence (stereo_rerender.py:720-725): rotate about the origin, then translate
    theta = 0.01
    R = mesh.get_rotation_matrix_from_xyz((0, -theta, 0))
    np.testing.assert_allclose(R, orc.rot_y(-theta), atol=1e-15)
    T = np.eye(4)
    T[:3, 3] = (0.1, 0.0, 0.2)
    mesh.transform(T)
    mesh.rotate(R, center=(0, 0, 0))
    mesh.translate([0.0315, 0.0, 0.0])
    np.testing.assert_allclose(mesh.pose, orc.eye_pose("left", 0.063, theta) @ T, atol=1e-15)
    np.testing.assert_allclose(mesh.vertices, orc.apply_pose(want, mesh.pose), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(mesh.get_center(), orc.apply_pose(want, mesh.pose).mean(axis=0), rtol=1e-11, atol=1e-12)
    # mesh reuse keeps the object, resets the pose
    mesh2, _ = dmt.get_mesh_from_depth_map(depth, K, colour, mesh, of_by_one=False)
    assert mesh2 is mesh and np.array_equal(mesh.pose, np.eye(4))


def test_render_matches_oracle_view():
    w, h = 160, 120
    depth_rgb, colour, depth = _frame(w, h, 1)
    K = dmt.compute_camera_matrix(60.0, None, w, h)
    mesh, _ = dmt.get_mesh_from_depth_map(depth, K, colour, None, of_by_one=False)
    mesh.translate([0.2, 0.0, 0.0])
    bg = np.array([0.0, 1.0, 0.0])
    img, zplane = dmt.render([mesh], K, depth=-2, bg_color=bg)
    assert img.dtype == np.float32 and img.shape == (h, w, 3) and zplane.shape == (h, w)
    M = np.eye(4)
    M[0, 3] = 0.2
    want, want_mask, ids = orc.render_view(depth_rgb, colour, 100, K, M, bg_rgb=(0, 255, 0), hole_fill=(0, 255, 0))
    got = (img * 255).astype(np.uint8)  # what the scripts do with render()'s result (stereo_rerender.py:819)
    u, v, z = orc.view_uvz(depth_rgb, 100, K, M)
    explained = explained_map(u, v, z, w, h)
    assert_differs_only_where_explained(got, want, explained)   # float32 render vs float64 oracle: rounding boundaries / z ties only
    holes = np.all(img == bg, axis=-1)   # stereo_rerender.py:740
    assert_differs_only_where_explained(holes, want_mask == 255, explained)
    assert holes.any()
    assert not ((np.abs(zplane - orc.zbuffer_depth(ids, z)) > 1e-4) & ~explained).any()   # rendered depth: another winner only where explained
    assert np.array_equal(dmt.render([mesh], K, depth=True), zplane)
    assert np.array_equal(dmt.render([mesh], K, bg_color=bg), img)


def test_render_mesh_plus_point_cloud_shares_one_zbuffer():
    w, h = 64, 48
    depth_rgb, colour, depth = _frame(w, h, 0, zero_fraction=0.0)
    K = dmt.compute_camera_matrix(60.0, None, w, h)
    mesh, _ = dmt.get_mesh_from_depth_map(depth, K, colour, None, of_by_one=False)
    # a red point in front of everything at the image centre, a blue one behind everything
    pts = np.array([[0.0, 0.0, 0.5], [0.0, 0.0, 50.0]])
    cols = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    pcd = dmt.pts_2_pcd(pts, cols)
    assert np.array_equal(pcd.points, pts) and np.array_equal(pcd.colors, cols)
    img = (dmt.render([mesh, pcd], K) * 255).astype(np.uint8)
    base = (dmt.render([mesh], K) * 255).astype(np.uint8)
    assert tuple(img[h // 2, w // 2]) == (255, 0, 0)
    diff = (img != base).any(axis=-1)
    assert diff.sum() == 1  # only the near point changed a pixel; the far one lost the depth test
    # convert_mesh_to_pcd drops the listed vertices from the render
    removed = np.arange(0, w * h, 2)
    pc = dmt.convert_mesh_to_pcd(mesh, removed, None)
    img2, zp = dmt.render([pc], K, depth=-2)
    ids = np.arange(w * h).reshape(h, w)
    assert (zp[ids % 2 == 0] == 0).all() and (zp[ids % 2 == 1] > 0).all()


# ---------------------------------------------------------------------------------------------
# reductions and export formats
# ---------------------------------------------------------------------------------------------
def test_depth_sums_and_centroid_vs_numpy():
    depth_rgb, _, _ = _frame(640, 480, 2)
    d3 = orc.decode_rgb_depth_frame(depth_rgb, 100, True, "D3").astype(np.float64)
    s = ops.depth_sums(cu(depth_rgb), 100, "D3").cpu().numpy()
    assert s[1] == d3.size and abs(s[0] - d3.sum()) <= 1e-12 * d3.sum() and abs(s[2] - (d3 * d3).sum()) <= 1e-12 * (d3 * d3).sum()
    mask = np.random.default_rng(0).integers(0, 256, (480, 640), dtype=np.uint8)
    sm = ops.depth_sums(cu(depth_rgb), 100, "D3", mask=cu(mask)).cpu().numpy()
    sel = d3[mask > 240]
    assert sm[1] == sel.size and abs(sm[0] - sel.sum()) <= 1e-12 * sel.sum()
    assert abs(sm[0] / sm[1] - orc.convergence_depth_of_frame(depth_rgb, 100, mask)) < 1e-5  # the reference's float32 mean
    empty = ops.depth_sums(cu(depth_rgb), 100, "D3", mask=torch.zeros((480, 640), dtype=torch.uint8, device=DEV)).cpu().numpy()
    assert empty[1] == 0 and empty[0] == 0
    a = ops.depth_sums(cu(depth_rgb), 100, "D3").cpu().numpy()
    assert np.array_equal(a, s)  # fixed summation order: bit-reproducible


def test_grey_and_touchly_exports():
    depth_rgb, _, _ = _frame(64, 48, 1)
    depth_rgb[..., 1] = np.random.default_rng(1).integers(0, 256, (48, 64), dtype=np.uint8)
    d2 = orc.decode_rgb_depth_frame(depth_rgb, 100, True, "D2")
    # convert_metric_depth_video_to_other_format.py:752-760
    want16 = np.rint(d2 * ((255 ** 2) / 100)).astype(np.uint16)
    want8 = np.repeat(np.rint(d2 * (255 / 100)).astype(np.uint8)[..., None], 3, axis=-1)
    assert np.array_equal(ops.depth_to_grey(cu(depth_rgb), 100, 16).cpu().numpy(), want16)
    assert np.array_equal(ops.depth_to_grey(cu(depth_rgb), 100, 8).cpu().numpy(), want8)
    # stereo_rerender.py:548-551 on the D1-decoded, master-FOV-scaled depth
    scale = orc.master_fov_depth_scale(45.0, 60.0)
    depth = orc.apply_depth_scale(orc.decode_rgb_depth_frame(depth_rgb, 100, True), scale)
    for tmin, tmax in ((0, 5), (0.5, 7.3)):
        d8 = np.rint(np.maximum(0, np.minimum(depth, tmax) - tmin) * (255 / (tmax - tmin))).astype(np.uint8)
        want = 255 - np.repeat(d8[..., None], 3, axis=-1)
        got = ops.touchly_depth(cu(depth_rgb), tmin, tmax, False, 100, "D1", scale).cpu().numpy()
        assert np.array_equal(got, want), (tmin, tmax)
        d8z = d8.copy()
        d8z[d8z == 0] = 255   # :688 / :827
        got_z = ops.touchly_depth(cu(depth), tmin, tmax, True, decoder="F32").cpu().numpy()
        assert np.array_equal(got_z, 255 - np.repeat(d8z[..., None], 3, axis=-1))
    stacked = torch.zeros((96, 64, 3), dtype=torch.uint8, device=DEV)
    ops.touchly_depth(cu(depth_rgb), 0, 5, False, 100, "D1", scale, out=stacked[48:])
    assert bool((stacked[:48] == 0).all()) and bool((stacked[48:] != 0).any())


# ---------------------------------------------------------------------------------------------
# novel view (3d_view_depthfile --render)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("of_by_one,yfov", [(True, None), (False, 50.0)])
def test_novel_view_renderer_vs_oracle(of_by_one, yfov):
    w, h, n = 160, 120, 2
    depth, colour = SyntheticClip(w, h, n, zero_fraction=0.005).frames()
    T = np.tile(np.eye(4), (n, 1, 1))
    T[1, :3, 3] = (0.1, 0.0, -0.2)
    nv = NovelViewRenderer(NovelViewParams(w, h, 60, yfov, 100, (2.0, 2.0, -4.0), (None, 0.5, None), T, of_by_one=of_by_one), DEV)
    rgb, mask = nv.render_device(cu(depth), cu(colour))
    for k in range(n):
        want, want_mask, ids, ext, uvz = orc.novel_view_frame(depth[k], colour[k], 60, yfov, 100, (2.0, 2.0, -4.0), (None, 0.5, None), T[k],
                                                              center_of_by_one=of_by_one, want_uvz=True)
        centre = nv.centroids(cu(depth[k:k + 1]), k)[0]
        np.testing.assert_allclose(nv.extrinsic(centre), ext, rtol=1e-9, atol=1e-9)
        explained = explained_map(*uvz, w, h)
        assert_differs_only_where_explained(rgb[k].cpu().numpy(), want, explained, 3e-3)
        assert_differs_only_where_explained(mask[k].cpu().numpy(), want_mask, explained, 3e-3)


def test_novel_view_render_host_pipeline_equals_render_device():
    """The three-stream host API of the novel view (uploads, kernels and downloads of neighbouring chunks overlap) returns
    complete buffers with exactly the frames of one device-resident call, for ragged chunkings and a start frame."""
    w, h, n = 96, 64, 7
    depth, colour = SyntheticClip(w, h, n, zero_fraction=0.005).frames()
    T = np.tile(np.eye(4), (n + 3, 1, 1))
    T[:, 0, 3] = np.linspace(0.0, 0.3, n + 3)
    nv = NovelViewRenderer(NovelViewParams(w, h, 60, None, 100, (2.0, 2.0, -4.0), (None, None, None), T), DEV)
    want, _ = nv.render_device(cu(depth), cu(colour), 3)
    want = want.cpu().numpy()
    for chunk in (1, 2, 3, 7, 16):
        got = nv.render_host(depth, colour, start_frame=3, chunk_frames=chunk)
        assert np.array_equal(np.asarray(got), want), chunk
    pinned = torch.empty((n, h, w, 3), dtype=torch.uint8, pin_memory=True)
    assert nv.render_host(torch.from_numpy(depth).pin_memory(), torch.from_numpy(colour).pin_memory(), pinned, start_frame=3) is pinned
    assert np.array_equal(pinned.numpy(), want)


@pytest.mark.parametrize("of_by_one,yfov,posed,size", [(True, None, True, (160, 120)), (False, 50.0, False, (160, 120)),
                                                       (True, 40.0, True, (160, 120)), (True, None, False, (70, 33)),
                                                       (False, None, True, (132, 50))])
def test_novel_view_device_camera_equals_host_camera(of_by_one, yfov, posed, size):
    """mdvt_novel_view_frames evaluates centroid -> look-at -> view on the device: its cameras must equal the NumPy
    helpers' (<= 1 float32 ulp per entry), and its images must be what the generic path renders from those cameras.
    (70, 33): widths that are not multiples of 4 take the scalar centroid / resolve kernels with the touched flags;
    (132, 50): rows that are not multiples of the 64-slot touched segments."""
    w, h = size
    n = 3
    depth, colour = SyntheticClip(w, h, n, zero_fraction=0.005).frames()
    T = None
    if posed:
        T = np.tile(np.eye(4), (n, 1, 1))
        T[1, :3, 3] = (0.1, 0.0, -0.2)
        c, s_ = np.cos(0.05), np.sin(0.05)
        T[2, :3, :3] = [[c, 0, s_], [0, 1, 0], [-s_, 0, c]]
    nv = NovelViewRenderer(NovelViewParams(w, h, 60, yfov, 100, (2.0, 2.0, -4.0), (None, 0.5, None), T, of_by_one=of_by_one), DEV)
    d, c_ = cu(depth), cu(colour)
    rgb, mask = nv.render_device(d, c_)
    views_dev = nv.last_views.cpu().numpy()
    sums_dev = nv.last_sums[:, :4].cpu().numpy()
    centres = nv.centroids(d)
    np.testing.assert_allclose(sums_dev[:, :3] / sums_dev[:, 3:4], centres, rtol=1e-13)
    host_views = []
    for k in range(n):
        v = nv.view_of(k, centres[k]).to_c()
        hv = np.array(list(v.M) + [v.fx, v.fy, v.cx, v.cy], dtype=np.float32)
        host_views.append(hv)
        ulps = np.abs(views_dev[k].view(np.int32).astype(np.int64) - hv.view(np.int32).astype(np.int64))
        tiny = np.abs(hv) < 1e-12   # +-0 entries
        assert (ulps[~tiny] <= 1).all(), (k, views_dev[k], hv)
        assert (np.abs(views_dev[k][tiny]) < 1e-12).all()
    # images: the generic path fed with the device's cameras, bit for bit
    specs = [[ops.ViewSpec(views_dev[k][:12].reshape(3, 4).astype(np.float64), *[float(x) for x in views_dev[k][12:]])] for k in range(n)]
    src = ops.make_source(w, h, nv.K, 100, "D1", True, 1.0, False)
    zb = ops.new_zbuf(1, w, h, DEV)
    rgb2 = torch.empty_like(rgb)
    mask2 = torch.empty_like(mask)
    ops.render_views(d, c_, [src], specs, w, h, zb, rgb2, mask2, None, (255, 255, 255), (255, 255, 255), 0)
    assert torch.equal(rgb, rgb2) and torch.equal(mask, mask2)
    assert bool((zb == -1).all())
    rgb3, mask3 = nv.render_device_hostcam(d, c_)
    assert (rgb3 != rgb).any(dim=-1).float().mean().item() < 1e-4
    assert bool((mask != 0).any()) and bool((mask == 0).any())


# ---------------------------------------------------------------------------------------------
# script front ends, end to end on small FFV1 clips
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def clip_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("clip")
    w, h, n = 128, 72, 7
    depth, colour = SyntheticClip(w, h, n, zero_fraction=0.005).frames()
    video_io.write_clip(str(d / "depth.mkv"), depth, 24.0)
    video_io.write_clip(str(d / "colour.mkv"), colour, 24.0)
    return dict(dir=d, depth=depth, colour=colour, w=w, h=h, n=n, depth_path=str(d / "depth.mkv"), colour_path=str(d / "colour.mkv"))


def test_cli_stereo_rerender_default_and_mask(clip_files):
    import stereo_rerender

    c = clip_files
    rc = stereo_rerender.main(["--depth_video", c["depth_path"], "--color_video", c["colour_path"], "--xfov", "60", "--infill_mask",
                               "--green_and_black_infill_mask", "--dont_place_points_in_edges", "--chunk_frames", "3"])
    assert rc == 0
    out = video_io.read_clip(c["depth_path"] + "_stereo.mkv")
    msk = video_io.read_clip(c["depth_path"] + "_stereo.mkv_infillmask.mkv")
    assert out.shape == (c["n"], c["h"], 2 * c["w"], 3) and msk.shape == out.shape
    assert not os.path.exists(c["depth_path"] + "_tmp_stereo.mkv")
    consts = ops.stereo_frame_constants(60.0, c["w"], 100, 63, 45.0)
    for k in range(c["n"]):
        want_sbs, want_mask, _ = km.stereo_rows_f32(c["depth"][k], c["colour"][k], consts, (0, 255, 0), (0, 0, 0), True)
        assert np.array_equal(out[k], want_sbs) and np.array_equal(msk[k], orc.mask_to_rgb(want_mask)), k
        ref_sbs, _, _ = orc.stereo_frame(c["depth"][k], c["colour"][k], 60.0, infill_mask=True)
        maps = [explained_map(*uvz, c["w"], c["h"]) for uvz in orc.stereo_frame_uvz(c["depth"][k], 60.0)]
        assert_differs_only_where_explained(out[k], ref_sbs, maps)


def test_cli_stereo_rerender_convergence_pose_and_max_frames(clip_files, tmp_path):
    import stereo_rerender

    c = clip_files
    conv = [4.0, float("nan"), 5.0, 6.0, 5.5, 5.0, 4.5]
    T = np.tile(np.eye(4), (c["n"], 1, 1))
    T[:, 0, 3] = np.linspace(0, 0.05, c["n"])
    json.dump(conv, open(tmp_path / "conv.json", "w"))
    json.dump(T.tolist(), open(tmp_path / "pose.json", "w"))
    json.dump([60.0 + k for k in range(c["n"])], open(tmp_path / "xfov.json", "w"))
    rc = stereo_rerender.main(["--depth_video", c["depth_path"], "--color_video", c["colour_path"], "--xfov_file", str(tmp_path / "xfov.json"),
                               "--convergence_file", str(tmp_path / "conv.json"), "--transformation_file", str(tmp_path / "pose.json"),
                               "--max_frames", "4"])
    assert rc == 0
    out = video_io.read_clip(c["depth_path"] + "_stereo.mkv")
    assert out.shape[0] == 4
    smooth = orc.smooth_convergence(orc.fill_nan_with_closest(conv))
    for k in range(4):
        ref, _, _ = orc.stereo_frame(c["depth"][k], c["colour"][k], 60.0 + k, convergence_depth=smooth[k], transform=T[k], infill_mask=False,
                                     tie_colour=True)
        maps = [explained_map(*uvz, c["w"], c["h"]) for uvz in orc.stereo_frame_uvz(c["depth"][k], 60.0 + k, convergence_depth=smooth[k], transform=T[k])]
        assert_differs_only_where_explained(out[k], ref, maps, 3e-3)


def test_cli_find_convergence_convert_and_view(clip_files, tmp_path):
    import importlib

    c = clip_files
    fcd = importlib.import_module("find_convergence_depth")
    assert fcd.main(["--depth_video", c["depth_path"]]) == 0
    got = json.load(open(c["depth_path"] + "_convergence_depths.json"))
    want = [orc.convergence_depth_of_frame(c["depth"][k]) for k in range(c["n"])]
    np.testing.assert_allclose(got, want, rtol=2e-6)

    conv = importlib.import_module("convert_metric_depth_video_to_other_format")
    ply_dir = tmp_path / "ply"
    assert conv.main(["--depth_video", c["depth_path"], "--color_video", c["colour_path"], "--xfov", "60", "--save_ply", str(ply_dir),
                      "--max_frames", "1", "--bit8"]) == 0
    files = sorted(os.listdir(ply_dir))
    assert files == ["0000000.ply", "0000001.ply", "0000002.ply"]  # the reference's loop converts max_frames + 2 frames (:763)
    for k, f in enumerate(files):
        xyz, rgb = orc.read_ply(str(ply_dir / f))
        want_xyz, want_rgb = orc.ply_points_of_frame(c["depth"][k], c["colour"][k], 60.0)
        assert np.array_equal(xyz.view(np.uint64), want_xyz.view(np.uint64)) and np.array_equal(rgb, want_rgb)
    grey = video_io.read_clip(c["depth_path"] + "_grey_depth.mkv")
    d2 = orc.decode_rgb_depth_frame(c["depth"][0], 100, True, "D2")
    assert np.array_equal(grey[0][..., 0], np.rint(d2 * (255 / 100)).astype(np.uint8))

    view = importlib.import_module("3d_view_depthfile")
    assert view.main(["--depth_video", c["depth_path"], "--color_video", c["colour_path"], "--xfov", "60", "--render", "--max_frames", "2"]) == 0
    out = video_io.read_clip(c["depth_path"] + "_render.mkv")
    assert out.shape == (2, c["h"], c["w"], 3)
    for k in range(2):
        want, _, _, _, uvz = orc.novel_view_frame(c["depth"][k], c["colour"][k], 60, center_of_by_one=True, tie_colour=True, want_uvz=True)
        assert_differs_only_where_explained(out[k], want, explained_map(*uvz, c["w"], c["h"]), 3e-3)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_cli_stereo_rerender_torchrun_two_ranks(clip_files, tmp_path):
    """Frames sharded over two ranks (NCCL broadcast of the parameter block, per-rank segments joined by rank 0)
    give the same file as the single-process run."""
    import shutil
    import subprocess
    import sys

    c = clip_files
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    work = tmp_path / "mg"
    work.mkdir()
    for name in ("depth.mkv", "colour.mkv"):
        shutil.copy(c["dir"] / name, work / name)
    conv = [4.0, 4.5, 5.0, 6.0, 5.5, 5.0, 4.5]
    json.dump(conv, open(work / "conv.json", "w"))
    base = ["--depth_video", str(work / "depth.mkv"), "--color_video", str(work / "colour.mkv"), "--xfov", "60", "--infill_mask",
            "--green_and_black_infill_mask", "--convergence_file", str(work / "conv.json")]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
           "29517", os.path.join(root, "stereo_rerender.py")] + base
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert proc.returncode == 0, proc.stderr[-2000:]
    sharded = video_io.read_clip(str(work / "depth.mkv") + "_stereo.mkv")
    sharded_mask = video_io.read_clip(str(work / "depth.mkv") + "_stereo.mkv_infillmask.mkv")
    assert not [f for f in os.listdir(work) if ".rank" in f]
    import stereo_rerender

    assert stereo_rerender.main(base) == 0
    single = video_io.read_clip(str(work / "depth.mkv") + "_stereo.mkv")
    assert sharded.shape == single.shape == (c["n"], c["h"], 2 * c["w"], 3)
    assert np.array_equal(sharded, single)
    assert np.array_equal(sharded_mask, video_io.read_clip(str(work / "depth.mkv") + "_stereo.mkv_infillmask.mkv"))


# ---------------------------------------------------------------------------------------------
# rendered depth planes, SBS depth video, Touchly1
# ---------------------------------------------------------------------------------------------
def _model_depth_planes(depth_rgb, consts, ids_pair):
    dec, scale = np.float32(consts[0]), np.float32(consts[1])
    c16 = (depth_rgb[..., 0].astype(np.uint32) << 8) | depth_rgb[..., 2].astype(np.uint32)
    z = (((c16 << 16).astype(np.float32) * dec) * scale).reshape(-1)
    return np.concatenate([np.where(ids >= 0, z[np.maximum(ids, 0)], np.float32(0)) for ids in ids_pair], axis=1).astype(np.float32)


@pytest.mark.parametrize("size,flags", [((640, 48), 0), ((70, 33), 0), ((1920, 8), 0x8)])  # fast kernel, any-width kernel, forced any-width
def test_stereo_rows_depth_plane_bit_exact(size, flags):
    w, h = size
    depth, colour = SyntheticClip(w, h, 2, zero_fraction=0.01).frames()
    consts = ops.stereo_frame_constants(60.0, w, 100, 63, 45.0)
    out_depth = torch.full((2, h, 2 * w), -1.0, dtype=torch.float32, device=DEV)
    ops.stereo_rows(cu(depth), cu(colour), cu(consts[None]), flags=flags, out_depth=out_depth)
    for f in range(2):
        _, _, ids = km.stereo_rows_f32(depth[f], colour[f], consts)
        assert np.array_equal(bits(out_depth[f].cpu().numpy()), bits(_model_depth_planes(depth[f], consts, ids))), f


def test_cli_sbs_depth_video_and_touchly1(clip_files, tmp_path):
    import shutil

    import stereo_rerender

    c = clip_files
    work = tmp_path / "modes"
    work.mkdir()
    for name in ("depth.mkv", "colour.mkv"):
        shutil.copy(c["dir"] / name, work / name)
    dv, cv = str(work / "depth.mkv"), str(work / "colour.mkv")
    consts = ops.stereo_frame_constants(60.0, c["w"], 100, 63, 45.0)
    # row-local path
    assert stereo_rerender.main(["--depth_video", dv, "--color_video", cv, "--xfov", "60", "--create_sbs_depth_video", "--max_frames", "3"]) == 0
    coded = video_io.read_clip(dv + "_stereo.mkv_depth.mkv")  # RGB order, 16-bit wire format
    sbs = video_io.read_clip(dv + "_stereo.mkv")
    assert coded.shape == sbs.shape == (3, c["h"], 2 * c["w"], 3)
    for k in range(3):
        want_sbs, _, ids = km.stereo_rows_f32(c["depth"][k], c["colour"][k], consts)
        assert np.array_equal(sbs[k], want_sbs)
        plane = _model_depth_planes(c["depth"][k], consts, ids)
        assert np.array_equal(coded[k], orc.encode_depth_frame_rgb(plane, 100, True)), k  # stereo_rerender.py:932-939
    # generic path (convergence): depth planes come from the 64-bit z-buffer
    json.dump([5.0] * c["n"], open(work / "conv.json", "w"))
    assert stereo_rerender.main(["--depth_video", dv, "--color_video", cv, "--xfov", "60", "--create_sbs_depth_video", "--max_frames", "2",
                                 "--convergence_file", str(work / "conv.json")]) == 0
    coded = video_io.read_clip(dv + "_stereo.mkv_depth.mkv")
    back = dfh.decode_rgb_depth_frame(coded[0], 100, True)
    K = orc.camera_matrix(60.0, None, c["w"], c["h"])
    scale = orc.master_fov_depth_scale(45.0, 60.0)
    theta = orc.convergence_angle(5.0 * scale, 0.063)
    for e, eye in enumerate(("left", "right")):
        u, v, z = orc.view_uvz(c["depth"][0], 100, K, orc.eye_pose(eye, 0.063, theta), depth_scale=scale)
        want = orc.zbuffer_depth(orc.splat_ids(u, v, z, c["w"], c["h"]), z)
        got = back[:, e * c["w"]:(e + 1) * c["w"]]
        off = np.abs(got - want) > 2e-3   # beyond the 1.55 mm quantisation of the wire format: another winner, i.e. a rounding boundary / z tie
        assert not (off & ~explained_map(u, v, z, c["w"], c["h"])).any() and off.mean() < 3e-3
    # touchly1 without a pose file: colour over reverse depth, no render (stereo_rerender.py:548-552)
    assert stereo_rerender.main(["--depth_video", dv, "--color_video", cv, "--xfov", "60", "--touchly1", "--touchly_max_depth", "7.5",
                                 "--max_frames", "2"]) == 0
    t1 = video_io.read_clip(dv + "_Touchly1.mkv")
    assert t1.shape == (2, 2 * c["h"], c["w"], 3)
    for k in range(2):
        depth = orc.apply_depth_scale(orc.decode_rgb_depth_frame(c["depth"][k], 100, True), scale)
        d8 = np.rint(np.maximum(0, np.minimum(depth, 7.5) - 0) * (255 / (7.5 - 0))).astype(np.uint8)
        assert np.array_equal(t1[k, :c["h"]], c["colour"][k]) and np.array_equal(t1[k, c["h"]:], 255 - np.repeat(d8[..., None], 3, axis=-1))
    # touchly1 with a pose file: mono render + rendered-depth plane (:677-692)
    T = np.tile(np.eye(4), (c["n"], 1, 1))
    T[:, 0, 3] = 0.04
    json.dump(T.tolist(), open(work / "pose.json", "w"))
    assert stereo_rerender.main(["--depth_video", dv, "--color_video", cv, "--xfov", "60", "--touchly1", "--max_frames", "1",
                                 "--transformation_file", str(work / "pose.json")]) == 0
    t1 = video_io.read_clip(dv + "_Touchly1.mkv")
    want_img, _, ids = orc.render_view(c["depth"][0], c["colour"][0], 100, K, T[0], depth_scale=scale)
    u, v, z = orc.view_uvz(c["depth"][0], 100, K, T[0], depth_scale=scale)
    assert_differs_only_where_explained(t1[0, :c["h"]], want_img, explained_map(u, v, z, c["w"], c["h"]), 3e-3)
    zplane = orc.zbuffer_depth(ids, z)
    d8 = np.rint(np.maximum(0, np.minimum(zplane, 5) - 0) * (255 / 5)).astype(np.uint8)
    d8[d8 == 0] = 255
    assert (t1[0, c["h"]:, :, 0] != 255 - d8).mean() < 5e-3


# ---------------------------------------------------------------------------------------------
# normals-coded infill mask (golden vectors made by running the reference, oracle/make_infill_golden.py)
# ---------------------------------------------------------------------------------------------
def _infill_case(golden_dir, tag):
    from oracle import infill_oracle as io

    g = np.load(os.path.join(golden_dir, "infill_mask.npz"))
    w, h, xfov, conv = g[tag + "_params"]
    w, h = int(w), int(h)
    K = orc.camera_matrix(xfov, None, w, h)
    scale = orc.master_fov_depth_scale(45.0, xfov)
    theta = None if conv == 0 else orc.convergence_angle(conv * scale, 0.063)
    M = orc.eye_pose("left", 0.063, theta) @ g[tag + "_transform"]
    return g, io, w, h, float(xfov), float(conv), K, scale, M


@pytest.mark.parametrize("tag", ["plain", "posed"])
def test_edge_vertices_match_reference_mesh_builder(golden_dir, tag):
    g, io, w, h, xfov, conv, K, scale, M = _infill_case(golden_dir, tag)
    src = ops.make_source(w, h, K, 100, "D1", True, scale, True)
    flags, normals = ops.edge_vertices(cu(g[tag + "_depth_rgb"]), src, K)
    idx = np.flatnonzero(flags.cpu().numpy().reshape(-1))
    assert np.array_equal(idx, g[tag + "_unused"])                       # the reference's unused_indices, exactly
    got = normals.cpu().numpy().reshape(-1, 3)[idx]
    np.testing.assert_allclose(got, g[tag + "_removed_normals"], rtol=0, atol=1e-14)
    # the drop-in module surface: get_mesh_from_depth_map(remove_edges=True, return_normals_of_removed=True)
    depth = orc.apply_depth_scale(orc.decode_rgb_depth_frame(g[tag + "_depth_rgb"], 100, True), scale)
    mesh, unused, removed = dmt.get_mesh_from_depth_map(depth, K, g[tag + "_colour"], None, remove_edges=True, of_by_one=True,
                                                        return_normals_of_removed=True)
    assert np.array_equal(unused, g[tag + "_unused"]) and np.abs(removed - g[tag + "_removed_normals"]).max() < 1e-14
    mesh2, used = dmt.get_mesh_from_depth_map(depth, K, g[tag + "_colour"], None, remove_edges=True, of_by_one=True)
    assert np.array_equal(np.setdiff1d(np.arange(w * h), used), g[tag + "_unused"])


@pytest.mark.parametrize("tag", ["plain", "posed"])
def test_infill_mask_painting_matches_reference_lines(golden_dir, tag):
    """E2 + E3 on the reference's own inputs: same left-eye image (holes = green) in, the reference's mask image before
    inpainting and its painted eye image out -- byte for byte; then the host finish against the reference's final mask."""
    from metric_depth_video_toolbox_b200 import infill

    g, io, w, h, xfov, conv, K, scale, M = _infill_case(golden_dir, tag)
    d, c = cu(g[tag + "_depth_rgb"]), cu(g[tag + "_colour"])
    src = ops.make_source(w, h, K, 100, "D1", True, scale, True)
    flags, normals = ops.edge_vertices(d, src, K)
    left = g[tag + "_left_image_u8"]
    hole = np.all(left == (0, 255, 0), axis=-1)
    image = left.copy()
    image[hole] = 0
    image_dev, hole_dev = cu(image), cu(hole.astype(np.uint8) * 255)
    mask_img = torch.zeros((h, w, 3), dtype=torch.uint8, device=DEV)
    zbuf = ops.new_zbuf(1, w, h, DEV)[0]
    ops.edge_splat(d, src, K, flags, M, K, w, h, zbuf)
    ops.edge_resolve(zbuf, d, src, K, normals, M, c, hole_dev, mask_img, image_dev)
    assert bool((zbuf == -1).all())
    assert np.array_equal(mask_img.cpu().numpy(), g[tag + "_mask_pre_inpaint"])
    assert np.array_equal(image_dev.cpu().numpy(), g[tag + "_image_final"])
    assert np.array_equal(infill.finish_mask(mask_img.cpu().numpy()), g[tag + "_mask_final"])
    # --green_and_black_infill_mask: no normals, no borders, but the edge colours are still painted (:777,794,813)
    image_dev2 = cu(image)
    ops.edge_splat(d, src, K, flags, M, K, w, h, zbuf)
    ops.edge_resolve(zbuf, d, src, K, None, M, c, hole_dev, mask_img, image_dev2, code_normals=False)
    assert np.array_equal(mask_img.cpu().numpy(), orc.mask_to_rgb(hole.astype(np.uint8) * 255))
    assert np.array_equal(image_dev2.cpu().numpy(), g[tag + "_image_final"])


def test_infill_mask_renderer_end_to_end(golden_dir):
    """InfillMaskRenderer (GPU render + E1-E3) against the oracle chain on a synthetic frame, both eyes."""
    from metric_depth_video_toolbox_b200 import infill
    from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
    from oracle import infill_oracle as io

    w, h = 160, 96
    depth, colour = SyntheticClip(w, h, 2, zero_fraction=0.003).frames()
    rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, infill_mask=True), DEV)
    im = infill.InfillMaskRenderer(rr, workers=2)
    sbs, mask_img = im.render_device(cu(depth), cu(colour))
    K = orc.camera_matrix(60.0, None, w, h)
    scale = orc.master_fov_depth_scale(45.0, 60.0)
    f = 1
    d_scaled = orc.apply_depth_scale(orc.decode_rgb_depth_frame(depth[f], 100, True), scale)
    unused, normals = io.edge_vertices(d_scaled, K, True)
    pts, ends = io.edge_points(d_scaled, K, unused, normals)
    for e, eye in enumerate(("left", "right")):
        M = orc.eye_pose(eye, 0.063, None)
        img, _, _ = orc.render_view(depth[f], colour[f], 100, K, M, depth_scale=scale, bg_rgb=(0, 255, 0), hole_fill=(0, 255, 0))
        want_mask, want_img, _, _ = io.mask_before_inpaint(img, colour[f], pts, ends, unused, M, K)
        got_mask = mask_img[f, :, e * w:(e + 1) * w].cpu().numpy()
        got_img = sbs[f, :, e * w:(e + 1) * w].cpu().numpy()
        assert (got_mask != want_mask).any(axis=-1).mean() < 3e-3 and (got_img != want_img).any(axis=-1).mean() < 3e-3
        assert (want_mask != 0).any()
    final = im.finish(mask_img.cpu().numpy())
    assert final.shape == (2, h, 2 * w, 3) and final.dtype == np.uint8
    assert np.array_equal(final[f, :, :w], infill.finish_mask(mask_img[f, :, :w].cpu().numpy()))


def test_cli_infill_mask_normals_coded(clip_files, tmp_path):
    """`--infill_mask` as movie_2_3D passes it (movie_2_3D.py:440): edge colours in the holes of the SBS video and the
    normals-coded, inpainted, blurred mask video -- against the oracle chain that is pinned to the reference's lines."""
    import shutil

    import stereo_rerender
    from metric_depth_video_toolbox_b200 import infill
    from oracle import infill_oracle as io

    c = clip_files
    work = tmp_path / "infill"
    work.mkdir()
    for name in ("depth.mkv", "colour.mkv"):
        shutil.copy(c["dir"] / name, work / name)
    dv = str(work / "depth.mkv")
    assert stereo_rerender.main(["--depth_video", dv, "--color_video", str(work / "colour.mkv"), "--xfov", "60", "--infill_mask",
                                 "--max_frames", "2"]) == 0
    sbs = video_io.read_clip(dv + "_stereo.mkv")
    msk = video_io.read_clip(dv + "_stereo.mkv_infillmask.mkv")
    w, h = c["w"], c["h"]
    assert sbs.shape == msk.shape == (2, h, 2 * w, 3)
    K = orc.camera_matrix(60.0, None, w, h)
    scale = orc.master_fov_depth_scale(45.0, 60.0)
    for k in range(2):
        d_scaled = orc.apply_depth_scale(orc.decode_rgb_depth_frame(c["depth"][k], 100, True), scale)
        unused, normals = io.edge_vertices(d_scaled, K, True)
        pts, ends = io.edge_points(d_scaled, K, unused, normals)
        for e, eye in enumerate(("left", "right")):
            M = orc.eye_pose(eye, 0.063, None)
            img, _, _ = orc.render_view(c["depth"][k], c["colour"][k], 100, K, M, depth_scale=scale, bg_rgb=(0, 255, 0), hole_fill=(0, 255, 0))
            pre, want_img, green, area = io.mask_before_inpaint(img, c["colour"][k], pts, ends, unused, M, K)
            assert (sbs[k, :, e * w:(e + 1) * w] != want_img).any(axis=-1).mean() < 3e-3
            # TELEA spreads a differing seed pixel over its neighbourhood: compare where the inputs agree
            want_final = io.finish_mask(pre, green, area)
            got_final = msk[k, :, e * w:(e + 1) * w]
            assert (np.abs(got_final.astype(int) - want_final.astype(int)) > 2).any(axis=-1).mean() < 0.02
            assert (want_final != 0).any()


def test_movie_2_3d_steps_4_and_5(clip_files, tmp_path):
    """BASELINE config 5 in miniature: precomputed depth + colour + focus-mask videos of one scene -> convergence list
    (step 4) -> stereo SBS + normals-coded infill mask videos (step 5), through the reference's step signatures."""
    import argparse
    import shutil

    import movie_2_3D
    from oracle import infill_oracle as io

    c = clip_files
    work = tmp_path / "movie"
    work.mkdir()
    for name in ("depth.mkv", "colour.mkv"):
        shutil.copy(c["dir"] / name, work / name)
    w, h, n = c["w"], c["h"], c["n"]
    focus = np.zeros((n, h, w, 3), dtype=np.uint8)
    focus[:, h // 4: h // 2, w // 4: w // 2] = 255
    # a speckled rim: the reference accepts a mask video only from 2 KB on (movie_2_3D.py:62-67), which a clean rectangle does not fill
    speckle = np.random.default_rng(3).integers(0, 2, size=(n, h, w), dtype=np.uint8) * 255
    focus[:, : h // 8] = speckle[:, : h // 8, :, None]
    video_io.write_clip(str(work / "mask.mkv"), focus, 24.0)
    scene = {"finished": False, "scene_video_file": str(work / "colour.mkv"), "depth_video_file": str(work / "depth.mkv"),
             "mask_video_file": str(work / "mask.mkv"), "xfov": 60.0, "sbs": str(work / "depth.mkv") + "_stereo.mkv"}
    movie_2_3D.step4_find_convergence([scene])
    conv = json.load(open(scene["convergence_file"]))
    want_conv = [orc.convergence_depth_of_frame(c["depth"][k], 100, focus[k, ..., 0]) for k in range(n)]
    np.testing.assert_allclose(conv, want_conv, rtol=2e-6)
    movie_2_3D.step5_render_sbs(argparse.Namespace(parallel=4), [scene])
    sbs = video_io.read_clip(scene["sbs"])
    msk = video_io.read_clip(scene["sbs"] + "_infillmask.mkv")
    assert sbs.shape == msk.shape == (n, h, 2 * w, 3)
    smooth = orc.smooth_convergence(orc.fill_nan_with_closest(conv))
    k = 3
    K = orc.camera_matrix(60.0, None, w, h)
    scale = orc.master_fov_depth_scale(45.0, 60.0)
    theta = orc.convergence_angle(smooth[k] * scale, 0.063)
    d_scaled = orc.apply_depth_scale(orc.decode_rgb_depth_frame(c["depth"][k], 100, True), scale)
    unused, normals = io.edge_vertices(d_scaled, K, True)
    pts, ends = io.edge_points(d_scaled, K, unused, normals)
    M = orc.eye_pose("left", 0.063, theta)
    img, _, _ = orc.render_view(c["depth"][k], c["colour"][k], 100, K, M, depth_scale=scale, bg_rgb=(0, 255, 0), hole_fill=(0, 255, 0))
    _, want_img, _, _ = io.mask_before_inpaint(img, c["colour"][k], pts, ends, unused, M, K)
    assert (sbs[k, :, :w] != want_img).any(axis=-1).mean() < 4e-3
    # a second call finds the outputs and skips the scene (movie_2_3D.py:431)
    movie_2_3D.step5_render_sbs(argparse.Namespace(parallel=4), [scene])


# ---------------------------------------------------------------------------------------------
# VR180: bit-exact cv2.remap on the GPU, --vr180 / --touchly0
# ---------------------------------------------------------------------------------------------
def test_remap_bilinear_bit_exact_vs_opencv(golden_dir):
    import cv2

    import stereo_rerender

    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (61, 83, 3), dtype=np.uint8)
    mx = rng.uniform(-4, 87, (70, 90)).astype(np.float32)
    my = rng.uniform(-4, 65, (70, 90)).astype(np.float32)
    mx[0, :5] = [0.0, 82.0, 82.5, -0.5, 1e9]          # exact pixel, last column, half out, half out, far out
    my[0, :5] = [0.0, 60.0, 60.5, -0.5, -1e9]
    mx[1, :4] = [10.015625, 10.484375, 10.5, 10.515625]  # ties of the 1/32-pixel quantisation (round half to even)
    my[1, :4] = 20.0
    want = cv2.remap(img, mx, my, interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=(0, 0, 0))
    got = ops.remap_bilinear(cu(img), cu(mx), cu(my)).cpu().numpy()
    assert np.array_equal(got, want)
    g = np.load(os.path.join(golden_dir, "misc.npz"))
    for k in range(3):  # the drop-in function against the reference's own outputs
        out = stereo_rerender.convert_to_equirectangular(g[f"equirect_in{k}"], input_fov=float(g[f"equirect_fov{k}"]))
        assert isinstance(out, np.ndarray) and np.array_equal(out, g[f"equirect_out{k}"]), k
    big = rng.integers(0, 256, (1920, 1920, 3), dtype=np.uint8)
    from metric_depth_video_toolbox_b200 import vr180

    bx, by = vr180.equirect_maps(1920, 1920, 75.0)
    want = cv2.remap(big, bx, by, interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=(0, 0, 0))
    assert np.array_equal(stereo_rerender.convert_to_equirectangular(big, input_fov=75.0), want)


def test_cli_vr180_and_touchly0(clip_files, tmp_path):
    import shutil

    import cv2

    import stereo_rerender
    from metric_depth_video_toolbox_b200 import vr180

    c = clip_files
    work = tmp_path / "vr"
    work.mkdir()
    for name in ("depth.mkv", "colour.mkv"):
        shutil.copy(c["dir"] / name, work / name)
    dv, cv = str(work / "depth.mkv"), str(work / "colour.mkv")
    assert stereo_rerender.main(["--depth_video", dv, "--color_video", cv, "--xfov", "60", "--touchly0", "--max_frames", "1"]) == 0
    out = video_io.read_clip(dv + "_Touchly0.mkv")
    side = 1920
    assert out.shape == (1, side, 3 * side, 3)
    w, h = c["w"], c["h"]
    K = orc.camera_matrix(60.0, None, w, h)
    fovx, fovy = orc.fov_of_camera_matrix(K)
    render_fov = max(75, max(fovx, fovy))
    Kr = orc.camera_matrix(render_fov, render_fov, side, side)
    scale = orc.master_fov_depth_scale(render_fov, 60.0)   # the render FOV is the master FOV in this mode (:534-538)
    mx, my = vr180.equirect_maps(side, side, render_fov)
    for e, eye in enumerate(("left", "right")):
        M = orc.eye_pose(eye, 0.063, None)
        img, _, ids = orc.render_view(c["depth"][0], c["colour"][0], 100, K, M, depth_scale=scale, K_out=Kr, out_size=(side, side))
        want = cv2.remap(img, mx, my, interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=(0, 0, 0))
        got = out[0, :, e * side:(e + 1) * side]
        assert (got != want).any(axis=-1).mean() < 5e-3 and want.any()
        if e == 0:  # third panel: the left eye's reverse depth through the same remap (:825-829,911-916)
            u, v, z = orc.view_uvz(c["depth"][0], 100, K, M, depth_scale=scale, K_out=Kr)
            zplane = orc.zbuffer_depth(ids, z)
            d8 = np.rint(np.maximum(0, np.minimum(zplane, 5) - 0) * (255 / 5)).astype(np.uint8)
            d8[d8 == 0] = 255
            panel = cv2.remap(np.repeat((255 - d8)[..., None], 3, axis=-1), mx, my, interpolation=cv2.INTER_LINEAR,
                              borderMode=cv2.BORDER_CONSTANT, borderValue=(0, 0, 0))
            assert (np.abs(out[0, :, 2 * side:].astype(int) - panel.astype(int)) > 1).any(axis=-1).mean() < 5e-3
    assert stereo_rerender.main(["--depth_video", dv, "--color_video", cv, "--xfov", "60", "--vr180", "--max_frames", "1"]) == 0
    assert video_io.read_clip(dv + "_stereo.mkv").shape == (1, side, 2 * side, 3)
    with pytest.raises(NotImplementedError, match="reference itself fails"):
        stereo_rerender.main(["--depth_video", dv, "--color_video", cv, "--xfov", "60", "--vr180", "--infill_mask"])


@pytest.mark.parametrize("tag", ["plain", "posed"])
def test_normal_march_infill_matches_reference(golden_dir, tag):
    """--do_basic_infill: the reference's infill_using_normals was run on its own mask / image; same bytes here."""
    g = np.load(os.path.join(golden_dir, "infill_mask.npz"))
    hole = g[tag + "_hole_mask"]
    image = g[tag + "_left_image_u8"].copy()
    image[hole] = 0
    dev = cu(image)
    ops.normal_march_infill(dev, cu(hole.astype(np.uint8) * 255), cu(g[tag + "_mask_final"]))
    assert np.array_equal(dev.cpu().numpy(), g[tag + "_image_basic_infill"])
    assert (g[tag + "_image_basic_infill"][hole] != 0).any()


def test_cli_do_basic_infill(clip_files, tmp_path):
    import shutil

    import stereo_rerender
    from oracle import infill_oracle as io

    c = clip_files
    work = tmp_path / "basic"
    work.mkdir()
    for name in ("depth.mkv", "colour.mkv"):
        shutil.copy(c["dir"] / name, work / name)
    dv = str(work / "depth.mkv")
    assert stereo_rerender.main(["--depth_video", dv, "--color_video", str(work / "colour.mkv"), "--xfov", "60", "--infill_mask",
                                 "--do_basic_infill", "--max_frames", "1"]) == 0
    sbs = video_io.read_clip(dv + "_stereo.mkv")
    msk = video_io.read_clip(dv + "_stereo.mkv_infillmask.mkv")
    w = c["w"]
    K = orc.camera_matrix(60.0, None, w, c["h"])
    scale = orc.master_fov_depth_scale(45.0, 60.0)
    M = orc.eye_pose("left", 0.063, None)
    img, hole_mask, _ = orc.render_view(c["depth"][0], c["colour"][0], 100, K, M, depth_scale=scale, bg_rgb=(0, 255, 0), hole_fill=(0, 0, 0))
    hole = hole_mask == 255
    # march with the mask the front end itself wrote: the image must then agree except at rounding-boundary pixels of the render
    want = io.normal_march_infill(img, hole, msk[0, :, :w])
    assert (sbs[0, :, :w] != want).any(axis=-1).mean() < 4e-3
    assert (want[hole] != 0).any(axis=-1).mean() > 0.5  # most holes got a colour
    with pytest.raises(NotImplementedError, match="--infill_mask"):
        stereo_rerender.main(["--depth_video", dv, "--xfov", "60", "--do_basic_infill"])
