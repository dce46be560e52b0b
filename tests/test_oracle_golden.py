"""Pin oracle/mdvt_oracle.py to vectors produced by RUNNING the reference (oracle/make_golden.py).

Bit-exact for the integer / float32 codec; <= 1e-12 relative for the float64 geometry; the
cv2.projectPoints twin (float32 K inside OpenCV) to <= 1e-4 px.
"""
import os

import numpy as np
import pytest

from oracle import mdvt_oracle as orc
from metric_depth_video_toolbox_b200.synth import SyntheticClip


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_decode_every_code_bit_exact(golden_dir):
    g = _load(golden_dir, "decode_all_codes.npz")
    rgb = g["rgb"]
    assert np.array_equal(orc.decode_codes(rgb, True, "D1"), g["d1_codes"])
    assert np.array_equal(orc.decode_codes(rgb, False, "D1"), g["d1_codes24"])
    for md in (100, 20):
        for variant, key in (("D1", "d1"), ("D2", "d2"), ("D3", "d3")):
            got = orc.decode_rgb_depth_frame(rgb, md, True, variant)
            ref = g[f"{key}_depth_md{md}"]
            assert got.dtype == np.float32 and ref.dtype == np.float32
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (variant, md)
        got24 = orc.decode_rgb_depth_frame(rgb, md, False, "D1")
        assert np.array_equal(got24.view(np.uint32), g[f"d1_depth24_md{md}"].view(np.uint32))


def test_known_answers():
    # fl32(100/255**4) = 0x1.964fcp-26 ; max code -> 101.57633 m ; D1 != D2/D3
    assert float(np.float32(100 / 255 ** 4)).hex() == "0x1.964fc00000000p-26"
    top = np.array([[[255, 255, 255]]], dtype=np.uint8)
    assert abs(float(orc.decode_rgb_depth_frame(top, 100)[0, 0]) - 101.57633) < 1e-4
    r, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    rgb = np.stack((r, r, b), axis=-1)
    d1 = orc.decode_rgb_depth_frame(rgb, 100, True, "D1")
    d3 = orc.decode_rgb_depth_frame(rgb, 100, True, "D3")
    assert int((d1.view(np.uint32) != d3.view(np.uint32)).sum()) == 48101
    K = orc.camera_matrix(60, None, 640, 480)
    assert abs(K[0, 0] - 554.25625842) < 1e-7 and K[1, 1] == K[0, 0] and K[0, 2] == 320 and K[1, 2] == 240


def test_encode_bit_exact(golden_dir):
    g = _load(golden_dir, "encode.npz")
    d = g["depth"]
    for md in (100, 20):
        codes = orc.encode_depth_codes(d, md)
        assert np.array_equal(codes, g[f"codes_md{md}"])
        assert np.array_equal(orc.codes_to_bgr(codes, True), g[f"bgr16_md{md}"])
        assert np.array_equal(orc.codes_to_bgr(codes, False), g[f"bgr24_md{md}"])


def test_wire_format_round_trip_error():
    rng = np.random.default_rng(3)
    d = rng.uniform(0, 100, (64, 64)).astype(np.float32)
    back = orc.decode_rgb_depth_frame(orc.encode_depth_frame_rgb(d, 100), 100)
    err = np.abs(back.astype(np.float64) - d)
    assert err.max() <= 1.56e-3  # "~1.5 mm" README.md:86-87
    assert np.all(back <= d + 1e-6)  # the encoder truncates


def test_camera_matrix(golden_dir):
    g = _load(golden_dir, "camera.npz")
    for case, K, fov in zip(g["cases"], g["K"], g["fov"]):
        fx = None if np.isnan(case[0]) else case[0]
        fy = None if np.isnan(case[1]) else case[1]
        got = orc.camera_matrix(fx, fy, int(case[2]), int(case[3]))
        assert np.array_equal(got, K)
        assert np.allclose(orc.fov_of_camera_matrix(got), fov, rtol=0, atol=1e-12)


@pytest.mark.parametrize("tag", ["64x48", "640x480"])
def test_geometry_stages(golden_dir, tag):
    g = _load(golden_dir, f"geometry_{tag}.npz")
    w, h = (int(x) for x in g["size"])
    stride = int(g["stride"])
    depth_rgb, colour = SyntheticClip(w, h, 3, seed=1234, zero_fraction=0.005).frame(1)
    if stride == 1:
        assert np.array_equal(depth_rgb, g["depth_rgb"]) and np.array_equal(colour, g["colour"])
    depth = orc.decode_rgb_depth_frame(depth_rgb, 100)
    assert np.array_equal(depth.reshape(-1)[::stride].view(np.uint32), g["depth_sample"].view(np.uint32))
    K = orc.camera_matrix(60.0, None, w, h)
    assert np.array_equal(K, g["K"])
    for obo in (0, 1):
        pts = orc.unproject(depth, K, bool(obo))
        assert pts.dtype == np.float64
        np.testing.assert_allclose(pts[::stride], g[f"xyz_obo{obo}"], rtol=1e-12, atol=0)
    pts = orc.unproject(depth, K, False)
    moved = orc.apply_pose(pts, g["T"])
    np.testing.assert_allclose(moved[::stride], g["xyz_T"], rtol=1e-12, atol=1e-13)
    u, v, z = orc.project(moved, K)
    front = (z > 0)[::stride]  # cv2.projectPoints maps z == 0 onto (cx, cy); the oracle culls instead
    ref = g["uv_T"]
    assert np.abs(u[::stride][front] - ref[front, 0]).max() < 1e-4
    assert np.abs(v[::stride][front] - ref[front, 1]).max() < 1e-4


def test_splat_matches_reference_point_painter(golden_dir):
    """ids -> image must equal the reference painter (np.round, in-bounds, argsort far->near,
    last write wins) wherever the nearest candidate is unique."""
    g = _load(golden_dir, "geometry_64x48.npz")
    w, h = 64, 48
    depth_rgb, colour = g["depth_rgb"], g["colour"]
    depth = orc.decode_rgb_depth_frame(depth_rgb, 100)
    K = g["K"]
    eye = orc.apply_pose(orc.unproject(depth, K, False), g["T"]) + np.array([0.0315, 0.0, 0.0])
    u, v, z = orc.project(eye, K)
    ids = orc.splat_ids(u, v, z, w, h)
    img, mask = orc.resolve(ids, colour, bg_rgb=(0, 0, 0), bg_collide=False)
    assert np.array_equal(mask == 0, g["painter_drawn"])
    # ties in z between candidates of one target are "undefined" in the reference (unstable
    # argsort); find them and exclude them from the exact comparison
    ur, vr = np.rint(u), np.rint(v)
    ok = (z > orc.NEAR_PLANE) & (ur >= 0) & (ur < w) & (vr >= 0) & (vr < h)
    tgt = (vr[ok] * w + ur[ok]).astype(np.int64)
    zz = z[ok]
    zmin = np.full(w * h, np.inf)
    np.minimum.at(zmin, tgt, zz)
    n_at_min = np.zeros(w * h, dtype=np.int64)
    np.add.at(n_at_min, tgt, (zz == zmin[tgt]).astype(np.int64))
    unique = (n_at_min <= 1).reshape(h, w)
    assert unique.mean() > 0.95
    assert np.array_equal(img[unique], g["painter_img"][unique])


def test_look_at_and_convergence(golden_dir):
    g = _load(golden_dir, "misc.npz")
    for k in range(3):
        pos, tgt = g[f"lookat_in{k}"]
        got = orc.look_at_extrinsic(np.asarray(pos, dtype=np.float32), tgt)
        np.testing.assert_allclose(got, g[f"lookat_out{k}"], rtol=1e-13, atol=1e-15)
    for (d, p), ref in zip(g["conv_angle_in"], g["conv_angle_out"]):
        assert orc.convergence_angle(d, p) == ref
    with pytest.raises(ValueError):
        orc.convergence_angle(0, 0.063)
    for k in range(4):
        filled = orc.fill_nan_with_closest(list(g[f"conv_in{k}"]))
        assert np.array_equal(np.array(filled), g[f"conv_filled{k}"])
        np.testing.assert_allclose(orc.smooth_convergence(filled), g[f"conv_smooth{k}"], rtol=1e-13, atol=0)
    with pytest.raises(ValueError):  # reference quirk: a 1-frame convergence list cannot be smoothed
        orc.smooth_convergence([3.0])


def test_eye_pose_algebra():
    ipd, th = 0.063, 0.01
    L, R = orc.eye_pose("left", ipd, th), orc.eye_pose("right", ipd, th)
    p = np.array([[0.3, -0.2, 4.0]])
    # left: Ry(-th) then +ipd/2 ; right: undo, Ry(+th) twice, -ipd/2 (stereo_rerender.py:723-725,831-836)
    left = (orc.rot_y(-th) @ p.T).T + [ipd / 2, 0, 0]
    right = (orc.rot_y(th) @ orc.rot_y(th) @ (left - [ipd / 2, 0, 0]).T).T - [ipd / 2, 0, 0]
    np.testing.assert_allclose(orc.apply_pose(p, L), left, rtol=1e-15)
    np.testing.assert_allclose(orc.apply_pose(p, R), right, rtol=1e-14)
    assert np.array_equal(orc.eye_pose("left", ipd, None)[:3, :3], np.eye(3))


def test_resolve_background_collision_quirk():
    ids = np.array([[0, 1, -1]])
    colour = np.array([[[0, 255, 0], [9, 8, 7], [1, 1, 1]]], dtype=np.uint8)
    img, mask = orc.resolve(ids, colour, bg_rgb=(0, 255, 0))
    assert mask.tolist() == [[255, 0, 255]]  # a rendered pure-green pixel reads as a hole
    assert img.tolist() == [[[0, 0, 0], [9, 8, 7], [0, 0, 0]]]
    assert orc.mask_to_rgb(mask).tolist() == [[[0, 255, 0], [0, 0, 0], [0, 255, 0]]]


def test_stereo_frame_identity_properties():
    """With ipd = 0 both eyes are the identity warp: every non-zero-depth pixel lands on itself."""
    depth_rgb, colour = SyntheticClip(64, 48, 2, zero_fraction=0.01).frame(0)
    sbs, mask, (idl, idr) = orc.stereo_frame(depth_rgb, colour, 60.0, pupillary_distance_mm=0, infill_mask=False)
    code = orc.decode_codes(depth_rgb)
    want = np.where(code.reshape(-1) > 0, np.arange(64 * 48), -1).reshape(48, 64)
    assert np.array_equal(idl, want) and np.array_equal(idr, want)
    assert sbs.shape == (48, 128, 3) and mask.shape == (48, 128)
    assert np.array_equal(sbs[:, :64][want >= 0], colour[want >= 0])


def test_ply_round_trip(tmp_path):
    depth_rgb, colour = SyntheticClip(64, 48, 1).frame(0)
    xyz, rgb = orc.ply_points_of_frame(depth_rgb, colour, 60.0)
    path = tmp_path / "0000000.ply"
    orc.write_ply(path, xyz, rgb)
    xyz2, rgb2 = orc.read_ply(path)
    assert np.array_equal(xyz, xyz2) and np.array_equal(rgb, rgb2)
    assert os.path.getsize(path) == len(xyz) * 27 + open(path, "rb").read().index(b"end_header\n") + 11
