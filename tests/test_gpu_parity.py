"""GPU parity tests proper: every call goes through the C ABI (ctypes -> libmdvt_b200.so).

  * integer / byte / index work: bit-exact against the golden vectors, the float64 oracle, or the
    float32 kernel model (which predicts the kernels' arithmetic bit for bit)
  * float work: <= 1e-4 relative against the float64 oracle (BASELINE.json north_star)
"""
import os

import numpy as np
import pytest
import torch

from oracle import kernel_model as km
from oracle import mdvt_oracle as orc
from metric_depth_video_toolbox_b200 import _lib, ops
from metric_depth_video_toolbox_b200.synth import SyntheticClip
from test_kernel_model import REL_TOL, boundary_explained, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ---------------------------------------------------------------------------------------------
# codec
# ---------------------------------------------------------------------------------------------
def test_decode_every_code_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "decode_all_codes.npz"))
    rgb = cu(g["rgb"])
    depth, codes = ops.decode_depth(rgb, 100, True, "D1", want_codes=True)
    assert np.array_equal(codes.cpu().numpy(), g["d1_codes"])
    codes24 = ops.decode_depth(rgb, 100, False, "D1", want_codes=True, want_depth=False)
    assert np.array_equal(codes24.cpu().numpy(), g["d1_codes24"])
    for md in (100, 20):
        for dec, key in (("D1", "d1"), ("D2", "d2"), ("D3", "d3")):
            got = ops.decode_depth(rgb, md, True, dec).cpu().numpy()
            assert np.array_equal(bits(got), bits(g[f"{key}_depth_md{md}"])), (dec, md)
        got24 = ops.decode_depth(rgb, md, False, "D1").cpu().numpy()
        assert np.array_equal(bits(got24), bits(g[f"d1_depth24_md{md}"]))


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 1023, 64 * 48 + 2])
def test_decode_ragged_lengths(n):
    rng = np.random.default_rng(n)
    rgb = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    for dec in ("D1", "D2", "D3"):
        depth, codes = ops.decode_depth(cu(rgb) if n else torch.empty((0, 3), dtype=torch.uint8, device=DEV), 100, True, dec,
                                        want_codes=True)
        want = orc.decode_rgb_depth_frame(rgb.reshape(n, 1, 3), 100, True, dec).reshape(n)
        assert np.array_equal(bits(depth.cpu().numpy()), bits(want))
        assert np.array_equal(codes.cpu().numpy(), orc.decode_codes(rgb.reshape(n, 1, 3), True, dec).reshape(n))


def test_decode_rejects_24bit_for_d2():
    with pytest.raises(_lib.MdvtError):
        ops.decode_depth(torch.zeros((4, 3), dtype=torch.uint8, device=DEV), 100, False, "D2")
    with pytest.raises(TypeError):
        ops.decode_depth(torch.zeros((4, 3), dtype=torch.uint8), 100)  # CPU tensor: no CPU path


def test_encode_golden_and_round_trip(golden_dir):
    g = np.load(os.path.join(golden_dir, "encode.npz"))
    d = cu(g["depth"])
    for md in (100, 20):
        pix16, codes = ops.encode_depth(d, md, True, True, want_codes=True)
        assert np.array_equal(codes.cpu().numpy(), g[f"codes_md{md}"])
        assert np.array_equal(pix16.cpu().numpy(), g[f"bgr16_md{md}"])
        assert np.array_equal(ops.encode_depth(d, md, False, True).cpu().numpy(), g[f"bgr24_md{md}"])
    rgb = ops.encode_depth(d, 100, True, False)
    back = ops.decode_depth(rgb, 100).cpu().numpy()
    clipped = np.clip(g["depth"], 0, 100)
    assert np.abs(back.astype(np.float64) - clipped).max() <= 1.56e-3


# ---------------------------------------------------------------------------------------------
# unprojection
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("of_by_one", [False, True])
@pytest.mark.parametrize("decoder", ["D1", "D2"])
def test_unproject_f64_bit_exact(of_by_one, decoder):
    w, h = 640, 480  # BASELINE config 1 size
    depth_rgb, _ = SyntheticClip(w, h, 2, zero_fraction=0.005).frame(0)
    depth_rgb[..., 1] = np.random.default_rng(0).integers(0, 256, (h, w), dtype=np.uint8)  # G differs from R: D2 != D1
    K = orc.camera_matrix(60.0, None, w, h)
    src = ops.make_source(w, h, K, 100, decoder, True, 1.0, of_by_one)
    got = ops.unproject(cu(depth_rgb), src, None, torch.float64, K).cpu().numpy()
    want = orc.unproject(orc.decode_rgb_depth_frame(depth_rgb, 100, True, decoder), K, of_by_one)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))  # bit-identical to NumPy
    T = np.eye(4)
    T[:3, :3] = orc.rot_y(0.02)
    T[:3, 3] = (0.1, 0.2, -0.3)
    got_t = ops.unproject(cu(depth_rgb), src, T, torch.float64, K).cpu().numpy()
    np.testing.assert_allclose(got_t, orc.apply_pose(want, T), rtol=1e-12, atol=1e-13)


def test_unproject_f32_model_and_tolerance():
    w, h = 640, 480
    depth_rgb, _ = SyntheticClip(w, h, 2, zero_fraction=0.005).frame(1)
    K = orc.camera_matrix(60.0, 47.0, w, h)
    scale = orc.master_fov_depth_scale(45.0, 60.0)
    src = ops.make_source(w, h, K, 100, "D1", True, scale, False)
    got = ops.unproject(cu(depth_rgb), src).cpu().numpy()
    X, Y, Z = km.unproject_f32(depth_rgb, km.source_constants(w, h, K, 100, "D1", scale))
    assert np.array_equal(bits(got[:, 0]), bits(X)) and np.array_equal(bits(got[:, 1]), bits(Y)) and np.array_equal(bits(got[:, 2]), bits(Z))
    want = orc.unproject(orc.apply_depth_scale(orc.decode_rgb_depth_frame(depth_rgb, 100), scale), K)
    assert rel_err(got, want, 1e-6).max() < REL_TOL


# ---------------------------------------------------------------------------------------------
# generic path: project + splat + resolve
# ---------------------------------------------------------------------------------------------
def _stereo_views(K, ipd, theta, T=None):
    T = np.eye(4) if T is None else T
    return [ops.ViewSpec(orc.eye_pose(e, ipd, theta) @ T, K[0, 0], K[1, 1], K[0, 2], K[1, 2]) for e in ("left", "right")]


@pytest.mark.parametrize("size,theta,posed", [((64, 48), None, False), ((640, 480), 0.008, False), ((640, 480), 0.008, True),
                                              ((70, 33), None, True)])
def test_project_splat_resolve(size, theta, posed):
    w, h = size
    depth_rgb, colour = SyntheticClip(w, h, 4, zero_fraction=0.005).frame(2)
    colour[1, 2] = (0, 255, 0)
    K = orc.camera_matrix(60.0, None, w, h)
    scale = orc.master_fov_depth_scale(45.0, 60.0)
    T = None
    if posed:
        T = np.eye(4)
        T[:3, :3] = orc.rot_y(0.01)
        T[:3, 3] = (0.05, -0.02, 0.1)
    views = _stereo_views(K, 0.063, theta, T)
    src = ops.make_source(w, h, K, 100, "D1", True, scale, False)
    zbuf = ops.new_zbuf(2, w, h, DEV)
    uvz = ops.project_splat(cu(depth_rgb), src, views, w, h, zbuf, want_uvz=True).cpu().numpy()
    msrc = km.source_constants(w, h, K, 100, "D1", scale)
    sbs = torch.zeros((h, 2 * w, 3), dtype=torch.uint8, device=DEV)
    msk = torch.zeros((h, 2 * w), dtype=torch.uint8, device=DEV)
    winners_differ = []
    for k, (eye, view) in enumerate(zip(("left", "right"), views)):
        # (1) float stage: bit-exact against the model, <= 1e-4 relative against the float64 oracle
        mu, mv, mz = km.view_uvz_f32(depth_rgb, msrc, view.M, (view.fx, view.fy, view.cx, view.cy))
        assert np.array_equal(bits(uvz[k, :, 2]), bits(mz))
        ok = mz > orc.NEAR_PLANE
        assert np.array_equal(bits(uvz[k, ok, 0]), bits(mu[ok])) and np.array_equal(bits(uvz[k, ok, 1]), bits(mv[ok]))
        u64, v64, z64 = orc.view_uvz(depth_rgb, 100, K, view.M, depth_scale=scale)
        ok64 = z64 > orc.NEAR_PLANE
        assert rel_err(uvz[k, ok64, 0], u64[ok64], 1.0).max() < REL_TOL
        assert rel_err(uvz[k, ok64, 1], v64[ok64], 1.0).max() < REL_TOL
        assert rel_err(uvz[k, ok64, 2], z64[ok64], 1e-6).max() < REL_TOL
        # (2) index stage: bit-exact given the same (u', v', z')
        want_ids = km.splat_ids_f32(uvz[k, :, 0], uvz[k, :, 1], uvz[k, :, 2], w, h)
        out_rgb, out_mask, depth_plane, ids = ops.resolve(zbuf[k], cu(colour), (0, 255, 0), (0, 0, 0), ops.FLAG_BG_COLLIDE | ops.FLAG_RESET_ZBUF,
                                                          out_rgb=sbs[:, k * w:(k + 1) * w], out_mask=msk[:, k * w:(k + 1) * w],
                                                          want_depth=True, want_ids=True)
        assert np.array_equal(ids.cpu().numpy().astype(np.int64), want_ids)
        img, mask = orc.resolve(want_ids, colour, (0, 255, 0), (0, 0, 0), True)
        assert np.array_equal(out_rgb.cpu().numpy(), img) and np.array_equal(out_mask.cpu().numpy(), mask)
        assert np.array_equal(bits(depth_plane.cpu().numpy()), bits(orc.zbuffer_depth(want_ids, uvz[k, :, 2])))
        # (3) end to end against the float64 oracle: only rounding-boundary / z-tie pixels may differ
        ids64 = orc.splat_ids(u64, v64, z64, w, h)
        n_diff, unexplained = boundary_explained(u64, v64, z64, want_ids, ids64, w, h)
        assert unexplained == 0 and n_diff <= max(4, int(2e-3 * w * h))
        winners_differ.append(want_ids != ids64)
    assert bool((zbuf == -1).all())  # FLAG_RESET_ZBUF left the buffer empty
    want_sbs, want_mask, _ = orc.stereo_frame(depth_rgb, colour, 60.0, convergence_depth=None if theta is None else 0.0315 / np.tan(theta) / scale,
                                              transform=T, infill_mask=True)
    # the final images differ from the float64 oracle's ONLY at the (explained, counted) pixels whose winner differs
    explained = np.concatenate(winners_differ, axis=1)
    assert not (((sbs.cpu().numpy() != want_sbs).any(axis=-1) | (msk.cpu().numpy() != want_mask)) & ~explained).any()


def test_project_splat_fast_division_equals_ieee_path():
    """The production splat shares one refined reciprocal per divisor; the parity-output variant (want_uvz) uses
    IEEE divisions.  Both must build the same z-buffer, bit for bit, also for a float32 depth source."""
    w, h = 640, 480
    depth_rgb, colour = SyntheticClip(w, h, 4, zero_fraction=0.01).frame(1)
    K = orc.camera_matrix(60.0, 41.0, w, h)
    T = np.eye(4)
    T[:3, :3] = orc.rot_y(0.05)
    T[:3, 3] = (0.3, -0.1, 0.2)
    views = _stereo_views(K, 0.063, 0.01, T)
    src = ops.make_source(w, h, K, 100, "D1", True, 1.37, False)
    za, zb, zc = (ops.new_zbuf(2, w, h, DEV) for _ in range(3))
    ops.project_splat(cu(depth_rgb), src, views, w, h, za, want_uvz=True)
    ops.project_splat(cu(depth_rgb), src, views, w, h, zb, want_uvz=False)
    assert torch.equal(za, zb) and bool((za != -1).any())
    depth = ops.decode_depth(cu(depth_rgb), 100)
    ops.project_splat(depth, ops.make_source(w, h, K, decoder="F32", depth_scale=1.37), views, w, h, zc)
    assert torch.equal(za, zc)


def test_resolve_mask_rgb_and_white_background():
    w, h = 64, 48
    depth_rgb, colour = SyntheticClip(w, h, 2).frame(0)
    K = orc.camera_matrix(60.0, None, w, h)
    views = _stereo_views(K, 0.3, None)[:1]
    src = ops.make_source(w, h, K, 100)
    zbuf = ops.new_zbuf(1, w, h, DEV)
    ops.project_splat(cu(depth_rgb), src, views, w, h, zbuf)
    rgb, mask3, _, ids = ops.resolve(zbuf[0], cu(colour), (255, 255, 255), (255, 255, 255), ops.FLAG_MASK_RGB, want_ids=True)
    ids = ids.cpu().numpy().astype(np.int64)
    img, mask = orc.resolve(ids, colour, (255, 255, 255), (255, 255, 255), False)
    assert np.array_equal(rgb.cpu().numpy(), img)
    assert np.array_equal(mask3.cpu().numpy(), orc.mask_to_rgb(mask, (255, 255, 255)))
    assert (mask == 255).any()


# ---------------------------------------------------------------------------------------------
# fused row-local stereo kernel
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [(64, 48), (640, 480), (70, 33), (48, 16), (1920, 8)])  # fast W%32, any-width, bulk W%16
@pytest.mark.parametrize("mask_rgb", [False, True])
def test_stereo_rows_bit_exact_vs_model(size, mask_rgb):
    w, h = size
    n = 3
    clip = SyntheticClip(w, h, n, zero_fraction=0.005)
    depth, colour = clip.frames()
    colour[:, 1, 2] = (0, 255, 0)
    xfovs = [60.0, 47.5, 75.0]
    consts = np.stack([ops.stereo_frame_constants(x, w, 100, 63, 45.0) for x in xfovs])
    flags = ops.FLAG_BG_COLLIDE | (ops.FLAG_MASK_RGB if mask_rgb else 0)
    sbs, mask = ops.stereo_rows(cu(depth), cu(colour), cu(consts), (0, 255, 0), (0, 0, 0), flags)
    sbs, mask = sbs.cpu().numpy(), mask.cpu().numpy()
    for f in range(n):
        want_sbs, want_mask, _ = km.stereo_rows_f32(depth[f], colour[f], consts[f], (0, 255, 0), (0, 0, 0), True)
        assert np.array_equal(sbs[f], want_sbs), f
        assert np.array_equal(mask[f], orc.mask_to_rgb(want_mask) if mask_rgb else want_mask), f
    # a single shared constants row == the same row repeated
    sbs1, mask1 = ops.stereo_rows(cu(depth), cu(colour), cu(consts[:1]), (0, 255, 0), (0, 0, 0), flags)
    sbsn, maskn = ops.stereo_rows(cu(depth), cu(colour), cu(np.repeat(consts[:1], n, 0)), (0, 255, 0), (0, 0, 0), flags)
    assert torch.equal(sbs1, sbsn) and torch.equal(mask1, maskn)


def test_stereo_rows_vs_float64_oracle():
    w, h = 640, 480
    depth_rgb, colour = SyntheticClip(w, h, 4, zero_fraction=0.005).frame(3)
    consts = ops.stereo_frame_constants(60.0, w, 100, 63, 45.0)
    sbs, mask = ops.stereo_rows(cu(depth_rgb[None]), cu(colour[None]), cu(consts[None]), (0, 255, 0), (0, 0, 0), ops.FLAG_BG_COLLIDE)
    want_sbs, want_mask, ids64 = orc.stereo_frame(depth_rgb, colour, 60.0, infill_mask=True)
    _, _, ids32 = km.stereo_rows_f32(depth_rgb, colour, consts, (0, 255, 0), (0, 0, 0), True)
    K = orc.camera_matrix(60.0, None, w, h)
    scale = orc.master_fov_depth_scale(45.0, 60.0)
    n_total = 0
    for eye, a, b in (("left", ids32[0], ids64[0]), ("right", ids32[1], ids64[1])):
        u64, v64, z64 = orc.view_uvz(depth_rgb, 100, K, orc.eye_pose(eye, 0.063, None), depth_scale=scale)
        n_diff, unexplained = boundary_explained(u64, v64, z64, a, b, w, h)
        assert unexplained == 0
        n_total += n_diff
    differing = (sbs[0].cpu().numpy() != want_sbs).any(axis=-1) | (mask[0].cpu().numpy() != want_mask)
    assert differing.sum() <= n_total <= max(4, int(2e-3 * w * h))


def test_stereo_rows_equals_generic_path_winners():
    """Same frame through K1+K2+K3 and through the fused kernel: images agree except where the two
    float32 formulas for u' straddle a rounding boundary."""
    w, h = 640, 480
    depth_rgb, colour = SyntheticClip(w, h, 4).frame(0)
    K = orc.camera_matrix(60.0, None, w, h)
    scale = orc.master_fov_depth_scale(45.0, 60.0)
    src = ops.make_source(w, h, K, 100, "D1", True, scale, False)
    zbuf = ops.new_zbuf(2, w, h, DEV)
    ops.project_splat(cu(depth_rgb), src, _stereo_views(K, 0.063, None), w, h, zbuf)
    generic = torch.cat([ops.resolve(zbuf[k], cu(colour))[0] for k in range(2)], dim=1)
    consts = ops.stereo_frame_constants(60.0, w, 100, 63, 45.0)
    fused, _ = ops.stereo_rows(cu(depth_rgb[None]), cu(colour[None]), cu(consts[None]))
    # every pixel where the two differ is a pixel where their float32 winners differ, and each of those is explained by a
    # float64 coordinate within 1e-3 px of a rounding boundary (the two kernels evaluate u' by different float32 formulas)
    _, _, ids_rows = km.stereo_rows_f32(depth_rgb, colour, consts)
    msrc = km.source_constants(w, h, K, 100, "D1", scale)
    differ = (generic != fused[0]).any(dim=-1).cpu().numpy()
    n_total = 0
    for k, eye in enumerate(("left", "right")):
        M = orc.eye_pose(eye, 0.063, None)
        ids_gen = km.splat_ids_f32(*km.view_uvz_f32(depth_rgb, msrc, M, (K[0, 0], K[1, 1], K[0, 2], K[1, 2])), w, h)
        u64, v64, z64 = orc.view_uvz(depth_rgb, 100, K, M, depth_scale=scale)
        n_diff, unexplained = boundary_explained(u64, v64, z64, ids_rows[k], ids_gen, w, h)
        assert unexplained == 0
        assert not (differ[:, k * w:(k + 1) * w] & ~(ids_rows[k] != ids_gen)).any()
        n_total += n_diff
    assert differ.sum() <= n_total <= max(4, int(2e-3 * w * h))


def test_stereo_rows_edge_cases():
    w, h = 64, 4
    consts = cu(ops.stereo_frame_constants(60.0, w, 100, 63, 45.0)[None])
    empty = torch.empty((0, h, w, 3), dtype=torch.uint8, device=DEV)
    sbs, mask = ops.stereo_rows(empty, empty, consts)
    assert sbs.shape == (0, h, 2 * w, 3) and mask.shape == (0, h, 2 * w)
    # all-zero depth: nothing is drawn, everything is a hole
    colour = cu(np.random.default_rng(0).integers(0, 256, (1, h, w, 3), dtype=np.uint8))
    sbs, mask = ops.stereo_rows(torch.zeros((1, h, w, 3), dtype=torch.uint8, device=DEV), colour, consts, fill_rgb=(7, 8, 9))
    assert bool((mask == 255).all()) and bool((sbs == torch.tensor([7, 8, 9], dtype=torch.uint8, device=DEV)).all())
    # maximum code everywhere: far plane, disparity < 0.5 px -> identity warp, no holes
    far = torch.full((1, h, w, 3), 255, dtype=torch.uint8, device=DEV)
    sbs, mask = ops.stereo_rows(far, colour, consts)
    assert bool((mask == 0).all()) and torch.equal(sbs[0, :, :w], colour[0]) and torch.equal(sbs[0, :, w:], colour[0])
    with pytest.raises(_lib.MdvtError):
        wide = torch.zeros((1, 1, 65536, 3), dtype=torch.uint8, device=DEV)
        ops.stereo_rows(wide, wide, consts)


def test_full_size_properties_1080p():
    """BASELINE configs[1] size.  Size-independent properties: (a) ipd = 0 is the identity warp;
    (b) the two eyes are mirror images of each other under a horizontal flip of the inputs;
    (c) every output pixel is either a hole or a colour present in the same source row."""
    w, h, n = 1920, 1080, 2
    clip = SyntheticClip(w, h, n, zero_fraction=0.005)
    depth, colour = clip.frames()
    d, c = cu(depth), cu(colour)
    zero_ipd = cu(ops.stereo_frame_constants(60.0, w, 100, 0, 45.0)[None])
    sbs, mask = ops.stereo_rows(d, c, zero_ipd)
    code_zero = torch.from_numpy((depth[..., 0] == 0) & (depth[..., 2] == 0)).to(DEV)
    for half in (slice(0, w), slice(w, 2 * w)):
        assert torch.equal(mask[:, :, half] == 255, code_zero)
        assert torch.equal(sbs[:, :, half][~code_zero], c[~code_zero])
    consts = cu(ops.stereo_frame_constants(60.0, w, 100, 63, 45.0)[None])
    sbs, mask = ops.stereo_rows(d, c, consts)
    # bit-exact against the model on one full-size frame
    want_sbs, want_mask, _ = km.stereo_rows_f32(depth[0], colour[0], ops.stereo_frame_constants(60.0, w, 100, 63, 45.0))
    assert np.array_equal(sbs[0].cpu().numpy(), want_sbs) and np.array_equal(mask[0].cpu().numpy(), want_mask)
    # (c) row-locality: sorted multiset check on a few rows via colour membership
    out = sbs[1].cpu().numpy()
    for r in (0, 517, 1079):
        row_cols = {tuple(px) for px in colour[1, r]}
        drawn = mask[1, r].cpu().numpy() == 0
        assert all(tuple(px) in row_cols for px in out[r][drawn][::37])
    hole_frac = (mask == 255).float().mean().item()
    assert 0.001 < hole_frac < 0.2


@pytest.mark.parametrize("threads", ["128", "160", "256", "320"])
def test_stereo_rows_every_cta_size_gives_identical_bytes(threads):
    """The fast kernel is instantiated for several CTA sizes (MDVT_ROW_THREADS); each must reproduce the model
    bit for bit.  The choice is read once per process, hence the subprocess."""
    import subprocess
    import sys

    code = (
        "import numpy as np, torch, sys\n"
        "sys.path.insert(0, %r)\n"
        "from metric_depth_video_toolbox_b200 import ops\n"
        "from metric_depth_video_toolbox_b200.synth import SyntheticClip\n"
        "from oracle import kernel_model as km\n"
        "for w, h in ((1920, 6), (640, 9), (96, 5), (3840, 3)):\n"
        "    d, c = SyntheticClip(w, h, 2, zero_fraction=0.01).frames()\n"
        "    c[:, 1, 2] = (0, 255, 0)\n"
        "    k = ops.stereo_frame_constants(60.0, w, 100, 63, 45.0)\n"
        "    sbs, m = ops.stereo_rows(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(k[None]).cuda(), (0, 255, 0), (0, 0, 0), ops.FLAG_BG_COLLIDE)\n"
        "    for f in range(2):\n"
        "        ws, wm, _ = km.stereo_rows_f32(d[f], c[f], k, (0, 255, 0), (0, 0, 0), True)\n"
        "        assert np.array_equal(sbs[f].cpu().numpy(), ws) and np.array_equal(m[f].cpu().numpy(), wm), (w, h, f)\n"
        "print('ok')\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MDVT_ROW_THREADS=threads)
    proc = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert proc.returncode == 0 and "ok" in proc.stdout, proc.stderr[-2000:]


def test_full_size_properties_4k_novel_view_and_rows():
    """BASELINE configs[2] size (3840x2160).  (a) a camera at the origin looking down +z is the identity warp: the
    render equals the colour frame wherever the depth code is non-zero, holes exactly where it is zero;
    (b) moving the camera keeps every drawn pixel a colour of the source frame and leaves holes;
    (c) the row kernel at 4K is bit-exact against the model on a band of rows and an ipd of 0 is the identity."""
    from metric_depth_video_toolbox_b200.novel_view import NovelViewParams, NovelViewRenderer

    w, h = 3840, 2160
    depth, colour = SyntheticClip(w, h, 1, zero_fraction=0.002).frames()
    d, c = cu(depth), cu(colour)
    nv = NovelViewRenderer(NovelViewParams(w, h, 60, None, 100, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0)), DEV)
    assert np.allclose(nv.extrinsic(np.zeros(3))[:3, :4], np.eye(4)[:3, :4])
    rgb, mask = nv.render_device(d, c)
    code_zero = torch.from_numpy((depth[0, ..., 0] == 0) & (depth[0, ..., 2] == 0)).to(DEV)
    assert torch.equal(mask[0] == 255, code_zero)
    assert torch.equal(rgb[0][~code_zero], c[0][~code_zero])
    white = torch.tensor([255, 255, 255], dtype=torch.uint8, device=DEV)
    assert bool((rgb[0][code_zero] == white).all())
    nv2 = NovelViewRenderer(NovelViewParams(w, h, 60, None, 100), DEV)  # the script's default camera (2, 2, -4) -> centroid
    rgb2, mask2 = nv2.render_device(d, c)
    hole_frac = (mask2 == 255).float().mean().item()
    assert 0.05 < hole_frac < 0.95
    drawn = (mask2[0] == 0)
    assert bool(((rgb2[0][~drawn] == white).all(dim=-1)).all())
    # (c) row kernel at W = 3840
    consts0 = cu(ops.stereo_frame_constants(60.0, w, 100, 0, 45.0)[None])
    sbs, m = ops.stereo_rows(d, c, consts0)
    for half in (slice(0, w), slice(w, 2 * w)):
        assert torch.equal(m[0, :, half] == 255, code_zero) and torch.equal(sbs[0, :, half][~code_zero], c[0][~code_zero])
    k = ops.stereo_frame_constants(60.0, w, 100, 63, 45.0)
    sbs, m = ops.stereo_rows(d, c, cu(k[None]))
    rows = slice(1000, 1016)
    want_sbs, want_mask, _ = km.stereo_rows_f32(depth[0, rows], colour[0, rows], k)
    assert np.array_equal(sbs[0, rows].cpu().numpy(), want_sbs) and np.array_equal(m[0, rows].cpu().numpy(), want_mask)


@pytest.mark.parametrize("size,conv,yfov,mask_rgb", [((64, 48), 5.0, None, False), ((640, 480), 2.0, None, True), ((640, 480), 0.4, 50.0, False),
                                                     ((70, 33), 3.0, None, False), ((1920, 24), 1.0, None, False)])
def test_stereo_conv_rows_bit_identical_to_generic_path(size, conv, yfov, mask_rgb):
    """The fused target-row kernel for convergence stereo against K1+K2+K3 with the same cameras: every output byte,
    the hole masks and the float32 depth planes must be identical (same float32 arithmetic, same winner order)."""
    from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer

    w, h = size
    n = 3
    depth, colour = SyntheticClip(w, h, n, zero_fraction=0.01).frames()
    colour[:, 1, 2] = (0, 255, 0)
    convs = [conv, 0.0, conv * 1.7]  # 0 -> "skipping convergence" for that frame (stereo_rerender.py:710-712)
    outs = []
    for force in (False, True):
        p = StereoParams(w, h, xfov=60.0, yfov=yfov, convergence_depths=convs, infill_mask=True, mask_rgb=mask_rgb, force_generic=force, conv_kernel=True)
        assert p.conv_local() != force
        rr = StereoRerenderer(p, DEV)
        out_depth = torch.full((n, h, 2 * w), -1.0, dtype=torch.float32, device=DEV)
        sbs, mask = rr.render_device(cu(depth), cu(colour), out_depth=out_depth)
        outs.append((sbs, mask, out_depth))
    (sa, ma, da), (sb, mb, db) = outs
    assert torch.equal(sa, sb) and torch.equal(ma, mb)
    assert torch.equal(da.view(torch.int32), db.view(torch.int32))
    assert bool((ma != 0).any()) and bool((da > 0).any())
    # both against the float32 model of the frame loops, every byte and every depth bit (frames 0 and 2: with / without rotation)
    scale = orc.master_fov_depth_scale(45.0, 60.0)
    K = orc.camera_matrix(60.0, yfov, w, h)
    msrc = km.source_constants(w, h, K, 100, "D1", scale)
    for f in (0, 1, 2):
        theta = None if convs[f] == 0 else orc.convergence_angle(convs[f] * scale, 0.063)
        views = [(orc.eye_pose(e, 0.063, theta), (K[0, 0], K[1, 1], K[0, 2], K[1, 2])) for e in ("left", "right")]
        img, mask, zplane, ids32 = km.render_views_f32(depth[f], colour[f], msrc, views, w, h, (0, 255, 0), (0, 0, 0), True)
        assert np.array_equal(sa[f].cpu().numpy(), img)
        got_mask = ma[f].cpu().numpy()
        assert np.array_equal(got_mask if not mask_rgb else (got_mask != 0).any(axis=-1).astype(np.uint8) * 255, mask)
        assert np.array_equal(bits(da[f].cpu().numpy()), bits(zplane))
    # and against the float64 reference restatement: winners differ only where a rounding boundary / z tie explains it
    _, _, ids64 = orc.stereo_frame(depth[2], colour[2], 60.0, yfov, convergence_depth=convs[2], infill_mask=True, tie_colour=True)
    n_total = 0
    for k, eye in enumerate(("left", "right")):
        u64, v64, z64 = orc.view_uvz(depth[2], 100, K, views[k][0], depth_scale=scale)
        n_diff, unexplained = boundary_explained(u64, v64, z64, ids32[k], ids64[k], w, h)
        assert unexplained == 0
        n_total += n_diff
    assert n_total <= max(4, int(3e-3 * w * h))


@pytest.mark.parametrize("size,convs,yfov,mask_rgb", [
    ((64, 48), [5.0, 0.0, 8.5], None, False),
    ((640, 480), [2.0, 0.0, 3.4], None, True),
    ((640, 480), [0.9, 1.3, 40.0], 50.0, False),
    ((1920, 24), [1.0, 0.0, 1.7], None, False),
    ((1920, 1080), [5.0, 1.1, 0.75], None, False),      # strong rotations: dozens of staircase steps per row, alternates beyond the staged slots
    ((3840, 2160), [4.0, 1.5], None, False),
    ((1280, 720), [3.0, 2.0, 1.0], None, True),
])
def test_stereo_conv_vrows_bit_identical_to_generic_path(size, convs, yfov, mask_rgb):
    """The virtual-source-row kernel (TMA-assembled rows, two 32-bit shared-memory atomic passes) against the generic frame
    loop with the same cameras: every output byte, the hole masks and the float32 depth planes are identical, no frame
    reports geometry outside its limits; small sizes also against the float32 model of the frame loops."""
    from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer

    w, h = size
    n = len(convs)
    depth, colour = SyntheticClip(w, h, n, zero_fraction=0.01).frames()
    colour[:, 1, 2] = (0, 255, 0)
    outs = []
    for kernel in ("vrows", "generic"):
        p = StereoParams(w, h, xfov=60.0, yfov=yfov, convergence_depths=convs, infill_mask=True, mask_rgb=mask_rgb, conv_kernel=kernel)
        rr = StereoRerenderer(p, DEV)
        out_depth = torch.full((n, h, 2 * w), -1.0, dtype=torch.float32, device=DEV)
        sbs, mask = rr.render_device(cu(depth), cu(colour), out_depth=out_depth)
        outs.append((sbs, mask, out_depth))
    (sa, ma, da), (sb, mb, db) = outs
    assert torch.equal(sa, sb) and torch.equal(ma, mb)
    assert torch.equal(da.view(torch.int32), db.view(torch.int32))
    assert bool((ma != 0).any()) and bool((da > 0).any())
    # the kernel's own geometry guard stayed silent, also when called through ops with a status plane
    rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, yfov=yfov, convergence_depths=convs, infill_mask=True, mask_rgb=mask_rgb), DEV)
    host = ops.conv_frames_packed(*rr.packed_cameras(0, n), rr.p.near)
    assert ops.conv_vrows_supported(host, w, h)
    status = torch.zeros(n, dtype=torch.int32, device=DEV)
    s2, m2 = ops.stereo_conv_rows(cu(depth), cu(colour), cu(host), rr.p.bg_rgb, (0, 0, 0),
                                  ops.FLAG_BG_COLLIDE | (ops.FLAG_MASK_RGB if mask_rgb else 0), kernel="vrows", status=status)
    assert int(status.sum().item()) == 0 and torch.equal(s2, sa) and torch.equal(m2.view(ma.shape), ma)
    if w * h <= 640 * 480:
        scale = orc.master_fov_depth_scale(45.0, 60.0)
        K = orc.camera_matrix(60.0, yfov, w, h)
        msrc = km.source_constants(w, h, K, 100, "D1", scale)
        for f in range(n):
            theta = None if convs[f] == 0 else orc.convergence_angle(convs[f] * scale, 0.063)
            views = [(orc.eye_pose(e, 0.063, theta), (K[0, 0], K[1, 1], K[0, 2], K[1, 2])) for e in ("left", "right")]
            img, mask, zplane, _ = km.render_views_f32(depth[f], colour[f], msrc, views, w, h, (0, 255, 0), (0, 0, 0), True)
            assert np.array_equal(sa[f].cpu().numpy(), img)
            got_mask = ma[f].cpu().numpy()
            assert np.array_equal(got_mask if not mask_rgb else (got_mask != 0).any(axis=-1).astype(np.uint8) * 255, mask)
            assert np.array_equal(bits(da[f].cpu().numpy()), bits(zplane))


def test_stereo_conv_vrows_without_mask_and_with_guard_columns():
    """No hole mask (infill_mask off: black background, no collision rule) and widths that do not fill the thread grid
    (96, 352, 1600: the guard instantiations) -- still byte for byte the generic loop."""
    from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer

    for (w, h), conv in (((96, 40), 2.0), ((352, 64), 1.5), ((1600, 48), 3.0)):
        depth, colour = SyntheticClip(w, h, 2, zero_fraction=0.02).frames()
        outs = []
        for kernel in ("vrows", "generic"):
            rr = StereoRerenderer(StereoParams(w, h, xfov=70.0, convergence_depths=[conv, conv * 2], infill_mask=False, conv_kernel=kernel), DEV)
            sbs, mask = rr.render_device(cu(depth), cu(colour))
            assert mask is None
            outs.append(sbs)
        assert torch.equal(outs[0], outs[1])
        assert bool((outs[0] != 0).any())


def test_stereo_conv_vrows_random_geometries_equal_the_generic_loop():
    """Thirty random geometries (frame size, field of view on both axes, pupillary distance, master FOV, per-frame convergence
    distances from 0.6 m to 60 m, background-collision on/off): wherever the host check accepts the poses, the virtual-row
    kernel writes the generic loop's bytes, masks and depth bits, and its own geometry guard stays silent."""
    from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer

    rng = np.random.default_rng(20260)
    taken = 0
    for case in range(30):
        w = 32 * int(rng.integers(1, 17))
        h = int(rng.integers(9, 120))
        n = int(rng.integers(1, 4))
        xfov = float(rng.uniform(35.0, 105.0))
        yfov = None if rng.random() < 0.5 else float(rng.uniform(30.0, 90.0))
        convs = [float(np.exp(rng.uniform(np.log(0.6), np.log(60.0)))) for _ in range(n)]
        ipd = float(rng.uniform(50.0, 75.0))
        infill = bool(rng.random() < 0.7)
        depth, colour = SyntheticClip(w, h, n, seed=100 + case, zero_fraction=0.01, n_rects=4).frames()
        common = dict(xfov=xfov, yfov=yfov, convergence_depths=convs, pupillary_distance=ipd, master_xfov=float(rng.uniform(40.0, 60.0)), infill_mask=infill)
        if rng.random() < 0.3:   # --xfov_file: a field of view per frame (the intrinsics change inside one launch)
            common.update(xfov=None, yfov=None, xfovs=[float(xfov + 3.0 * k) for k in range(n)])
        probe = StereoRerenderer(StereoParams(w, h, **common), DEV)
        host = ops.conv_frames_packed(*probe.packed_cameras(0, n), probe.p.near)
        if not ops.conv_vrows_supported(host, w, h):
            continue
        taken += 1
        outs = []
        for kernel in ("vrows", "generic"):
            rr = StereoRerenderer(StereoParams(w, h, conv_kernel=kernel, **common), DEV)
            out_depth = torch.full((n, h, 2 * w), -1.0, dtype=torch.float32, device=DEV)
            sbs, mask = rr.render_device(cu(depth), cu(colour), out_depth=out_depth)
            outs.append((sbs, mask, out_depth))
        (sa, ma, da), (sb, mb, db) = outs
        assert torch.equal(sa, sb), (case, w, h, xfov, yfov, convs)
        assert (ma is None and mb is None) or torch.equal(ma, mb), (case, w, h)
        assert torch.equal(da.view(torch.int32), db.view(torch.int32)), (case, w, h)
        status = torch.zeros(n, dtype=torch.int32, device=DEV)
        ops.stereo_conv_rows(cu(depth), cu(colour), cu(host), probe.p.bg_rgb, (0, 0, 0), ops.FLAG_BG_COLLIDE if infill else 0, kernel="vrows",
                             status=status, want_mask=infill)
        assert int(status.sum().item()) == 0
    assert taken >= 20


def test_pose_file_of_pans_takes_the_virtual_row_kernel_and_other_poses_the_generic_loop():
    """--transformation_file: a camera that only pans (rotation about y) keeps every eye pose a `y rotation + x shift`, so "auto"
    renders it with the virtual-row kernel; one frame with a tilt -- or any translation, which the convergence rotation turns
    into a shift along z -- sends the chunk through the generic loop.  Either way the bytes are the generic loop's."""
    from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer

    w, h, n = 256, 72, 3
    depth, colour = SyntheticClip(w, h, n, zero_fraction=0.01).frames()

    def pan(angle, tx):
        T = np.eye(4)
        T[:3, :3] = orc.rot_y(angle)
        T[0, 3] = tx
        return T

    pans = [pan(0.004 * k, 0.0) for k in range(n)]
    slides = [pan(0.004 * k, 0.03 * k) for k in range(n)]
    tilted = [T.copy() for T in pans]
    tilted[1][:3, :3] = tilted[1][:3, :3] @ np.array([[1, 0, 0], [0, np.cos(0.002), -np.sin(0.002)], [0, np.sin(0.002), np.cos(0.002)]])
    for poses, expect_vrows in ((pans, True), (tilted, False), (slides, False)):
        outs = []
        for kernel in ("auto", "generic"):
            rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=[3.0] * n, transformations=poses, conv_kernel=kernel), DEV)
            sbs, mask = rr.render_device(cu(depth), cu(colour))
            if kernel == "auto":
                assert rr._consts_cache[("conv", 0, n)][1] == expect_vrows
            outs.append((sbs, mask))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_stereo_conv_vrows_limits_send_extreme_poses_to_the_generic_loop():
    """Outside the virtual-row kernel's limits (width not a multiple of 32, convergence so close that the staircase is steeper
    than 0.4 rows per 15 columns) "auto" renders through the generic loop -- same bytes as asking for it -- and "vrows" refuses."""
    from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer

    for (w, h), conv in (((70, 33), 3.0), ((640, 480), 0.05)):
        depth, colour = SyntheticClip(w, h, 2).frames()
        outs = []
        for kernel in ("auto", "generic"):
            rr = StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=[conv, conv], conv_kernel=kernel), DEV)
            outs.append(rr.render_device(cu(depth), cu(colour)))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        host = ops.conv_frames_packed(*rr.packed_cameras(0, 2), rr.p.near)
        assert not ops.conv_vrows_supported(host, w, h)
        with pytest.raises(ValueError):
            StereoRerenderer(StereoParams(w, h, xfov=60.0, convergence_depths=[conv, conv], conv_kernel="vrows"), DEV).render_device(cu(depth), cu(colour))


def test_pack_mask_bits_equals_numpy_packbits_and_host_api():
    """One bit per mask pixel (what render_host(mask_format="bits") ships over PCIe): np.packbits order, any non-zero byte
    counts as set, ragged tails, and the host API's packed masks unpack to exactly its u8 masks."""
    from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer

    rng = np.random.default_rng(3)
    for shape in ((7, 64), (3, 5, 3840), (1, 8), (2, 40)):
        m = (rng.random(shape) < 0.3).astype(np.uint8) * 255
        m[..., ::7] = np.where(m[..., ::7] == 0, 0, rng.integers(1, 256, size=m[..., ::7].shape)).astype(np.uint8)  # non-zero, not 255
        got = ops.pack_mask_bits(cu(m)).cpu().numpy()
        assert np.array_equal(got, np.packbits(m != 0, axis=-1))
        assert np.array_equal(ops.unpack_mask_bits(got), (m != 0).astype(np.uint8) * 255)
    w, h, n = 640, 96, 5
    depth, colour = SyntheticClip(w, h, n, zero_fraction=0.01).frames()
    rr = StereoRerenderer(StereoParams(w, h, xfov=60.0), DEV)
    sbs_a, mask_u8 = rr.render_host(depth, colour, chunk_frames=2)
    sbs_b, mask_bits = rr.render_host(depth, colour, chunk_frames=2, mask_format="bits")
    assert torch.equal(sbs_a, sbs_b) and tuple(mask_bits.shape) == (n, h, 2 * w // 8)
    assert np.array_equal(ops.unpack_mask_bits(mask_bits), mask_u8.numpy()) and bool((mask_u8 != 0).any())
