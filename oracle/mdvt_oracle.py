"""CPU oracle for the dense per-frame path (decode -> unproject -> pose -> project -> splat).

TEST INFRASTRUCTURE ONLY.  Nothing under ``metric_depth_video_toolbox_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` do, and there only as the checker / the timed CPU baseline.

This is a from-scratch NumPy restatement of what calledit/metric_depth_video_toolbox computes
per pixel per frame.  Every function cites the reference file:line it follows (paths relative
to the reference checkout).  Arithmetic types follow the reference under NumPy >= 2 promotion
rules: the depth decode is uint32 -> one float32 multiply (or divide), everything geometric
is float64.

Parity pin: the reference ships no tests or golden vectors.  ``oracle/make_golden.py`` runs the
reference's own importable functions (and exec()s its inline decode / point-painter lines
straight from the reference files) on seeded inputs and stores the results in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` holds this module to those vectors
(bit-exact for integer / float32 work, <= 1e-12 relative for float64 geometry).  The rasteriser
(`depth_map_tools.render`, Open3D/OpenGL, Windows-only in the pinned snapshot) cannot run
anywhere on Linux, so image parity is pinned on the reference's *point painter* rule
(stereo_rerender.py:746-755,814): round-half-even target pixel, nearest z wins.
"""
from __future__ import annotations

import math

import numpy as np

# 255**4: the full-scale code of the wire format (depth_frames_helper.py:8,22)
FULL_SCALE = 255 ** 4
NEAR_PLANE = 1e-4  # depth_map_tools.py:1520 (set_constant_z_near(0.0001))

DECODER_VARIANTS = ("D1", "D2", "D3")


# --------------------------------------------------------------------------------------
# wire-format codec
# --------------------------------------------------------------------------------------
def decode_codes(rgb: np.ndarray, bit16: bool = True, variant: str = "D1") -> np.ndarray:
    """RGB-order u8 frame -> uint32 depth code.

    D1  depth_frames_helper.py:63-75  (bit16: byte3 <- R, byte2 <- B, G ignored;
                                       24-bit: byte0 <- B, byte1 <- R, byte2 <- G -- the
                                       reference's channel order, not the encoder's)
    D2  convert_metric_depth_video_to_other_format.py:646-649  (byte3 <- trunc((R+G)/2))
    D3  find_convergence_depth.py:56-59                         (byte3 <- R)
    """
    r = rgb[..., 0].astype(np.uint32)
    g = rgb[..., 1].astype(np.uint32)
    b = rgb[..., 2].astype(np.uint32)
    if variant == "D1":
        if bit16:
            return (r << 24) | (b << 16)
        return b | (r << 8) | (g << 16)
    if variant == "D2":
        hi = (r + g) >> 1  # float64 /2 then truncating cast to u8 == integer halving
        return (hi << 24) | (b << 16)
    if variant == "D3":
        return (r << 24) | (b << 16)
    raise ValueError(f"unknown decoder variant {variant!r}")


def codes_to_depth(codes: np.ndarray, max_depth, variant: str = "D1") -> np.ndarray:
    """uint32 code -> float32 metres.

    D1 multiplies by fl32(max_depth / 255**4)   (depth_frames_helper.py:13-24)
    D2/D3 divide by fl32(255**4 / max_depth)    (convert_...py:652, find_convergence_depth.py:60)
    Both are a single IEEE float32 operation on fl32(code).
    """
    e = codes.astype(np.float32)
    if variant == "D1":
        return e * np.float32(float(max_depth) / FULL_SCALE)
    return e / np.float32(FULL_SCALE / max_depth)


def decode_rgb_depth_frame(rgb, max_depth, bit16=True, variant="D1"):
    """depth_frames_helper.py:99-103 (D1) and its two inline cousins (D2, D3)."""
    return codes_to_depth(decode_codes(rgb, bit16, variant), max_depth, variant)


def encode_depth_codes(depth: np.ndarray, max_depth) -> np.ndarray:
    """float depth -> uint32 code, truncating.  depth_frames_helper.py:5-11."""
    d = np.clip(depth, 0.0, max_depth).astype(np.float64)
    return ((FULL_SCALE / float(max_depth)) * d).astype(np.uint32)


def codes_to_bgr(codes: np.ndarray, bit16: bool = False) -> np.ndarray:
    """uint32 code -> (H,W,3) u8 in B,G,R order for cv2.  depth_frames_helper.py:48-61."""
    c = codes.astype(np.uint32)
    if bit16:
        hi = ((c >> 24) & 0xFF).astype(np.uint8)
        lo = ((c >> 16) & 0xFF).astype(np.uint8)
        return np.stack((lo, hi, hi), axis=-1)
    b0 = (c & 0xFF).astype(np.uint8)
    b1 = ((c >> 8) & 0xFF).astype(np.uint8)
    b2 = ((c >> 16) & 0xFF).astype(np.uint8)
    return np.stack((b0, b1, b2), axis=-1)


def encode_depth_frame_rgb(depth, max_depth, bit16=True):
    """Convenience: metres -> RGB-order frame as the scripts see it after BGR2RGB."""
    return np.ascontiguousarray(codes_to_bgr(encode_depth_codes(depth, max_depth), bit16)[..., ::-1])


# --------------------------------------------------------------------------------------
# camera model
# --------------------------------------------------------------------------------------
def camera_matrix(fov_x_deg, fov_y_deg, width, height) -> np.ndarray:
    """depth_map_tools.py:902-934.  A missing FOV copies the other axis' focal length."""
    fx = fy = None
    if fov_x_deg is not None:
        fx = width / (2 * np.tan(np.deg2rad(fov_x_deg) / 2))
    if fov_y_deg is not None:
        fy = height / (2 * np.tan(np.deg2rad(fov_y_deg) / 2))
    if fy is None:
        fy = fx
    if fx is None:
        fx = fy
    return np.array([[fx, 0, width / 2], [0, fy, height / 2], [0, 0, 1]], dtype=np.float64)


def fov_of_camera_matrix(K):
    """depth_map_tools.py:1640-1649."""
    w, h = K[0][2] * 2, K[1][2] * 2
    return (np.rad2deg(2 * np.arctan2(w, 2 * K[0][0])), np.rad2deg(2 * np.arctan2(h, 2 * K[1][1])))


def master_fov_depth_scale(master_xfov_deg, xfov_deg) -> float:
    """stereo_rerender.py:537-538: 1 / (tan(master/2) / tan(xfov/2)), Python floats."""
    return 1.0 / (math.tan(math.radians(master_xfov_deg / 2)) / math.tan(math.radians(xfov_deg / 2)))


def apply_depth_scale(depth_f32: np.ndarray, scale: float) -> np.ndarray:
    """stereo_rerender.py:541: in-place float32 *= Python float -> one float32 multiply."""
    return depth_f32 * np.float32(scale)


# --------------------------------------------------------------------------------------
# unprojection / pose / projection  (float64)
# --------------------------------------------------------------------------------------
def unproject(depth: np.ndarray, K: np.ndarray, of_by_one: bool = False) -> np.ndarray:
    """depth (H,W) -> (H*W,3) float64 points, row-major.  depth_map_tools.py:1112-1133.

    of_by_one stretches the pixel grid by (W+1)/W, (H+1)/H in float32 first (:1118-1123).
    Expression order is (x - cx) * z / fx, left to right.
    """
    h, w = depth.shape
    jj, ii = np.meshgrid(np.arange(w), np.arange(h))
    if of_by_one:
        jj = jj.astype(np.float32) * np.float32((w + 1) / w)
        ii = ii.astype(np.float32) * np.float32((h + 1) / h)
    z = depth
    x3 = (jj - np.float64(K[0][2])) * z / np.float64(K[0][0])
    y3 = (ii - np.float64(K[1][2])) * z / np.float64(K[1][1])
    out = np.empty((h * w, 3), dtype=np.float64)
    out[:, 0] = x3.reshape(-1)
    out[:, 1] = y3.reshape(-1)
    out[:, 2] = np.asarray(z, dtype=np.float64).reshape(-1)
    return out


def apply_pose(points: np.ndarray, T: np.ndarray) -> np.ndarray:
    """4x4 affine on (N,3) points; w is dropped, not divided.  depth_map_tools.py:977-1004
    (Open3D's mesh.transform, stereo_rerender.py:616, agrees for affine T)."""
    T = np.asarray(T, dtype=np.float64)
    return points @ T[:3, :3].T + T[:3, 3]


def rot_y(angle: float) -> np.ndarray:
    """Open3D get_rotation_matrix_from_xyz((0, a, 0)) == Ry(a).  stereo_rerender.py:719-720."""
    c, s = math.cos(angle), math.sin(angle)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=np.float64)


def convergence_angle(distance: float, pupillary_distance: float) -> float:
    """stereo_rerender.py:94-112."""
    if distance == 0:
        raise ValueError("Distance must be non-zero to compute a valid angle.")
    return math.atan((pupillary_distance / 2) / distance)


def eye_pose(eye: str, ipd_m: float, conv_angle) -> np.ndarray:
    """4x4 that takes frame-space points into the eye camera.

    LEFT : rotate Ry(-theta) about the origin, then translate +ipd/2   (stereo_rerender.py:723-725)
    RIGHT: net Ry(+theta), translate -ipd/2                            (stereo_rerender.py:831-836)
    theta is None / 0 without a convergence file (:708-721).
    """
    sign = {"left": +1.0, "right": -1.0}[eye]
    M = np.eye(4)
    if conv_angle:
        M[:3, :3] = rot_y(-sign * conv_angle)
    M[0, 3] = sign * ipd_m / 2
    return M


def project(points: np.ndarray, K: np.ndarray):
    """Pinhole projection u = fx X/Z + cx, v = fy Y/Z + cy (depth_map_tools.py:1057-1060 is the
    cv2 twin; 1523-1552 is the GL one).  No culling here; returns (u, v, z)."""
    z = points[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = K[0][0] * points[:, 0] / z + K[0][2]
        v = K[1][1] * points[:, 1] / z + K[1][2]
    return u, v, z


def look_at_extrinsic(cam_pos, target, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """The reference's non-standard look-at (depth_map_tools.py:1618-1638): r,u,f as *columns*
    of the upper 3x3, translation (px, py, -pz).  Open3D consumes the upper 3x4 as
    world->camera."""
    cam_pos = np.asarray(cam_pos, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    f = target - cam_pos
    f = f / np.linalg.norm(f)
    r = np.cross(np.asarray(up, dtype=np.float64), f)
    r = r / np.linalg.norm(r)
    u = np.cross(f, r)
    M = np.eye(4)
    M[:3, 0], M[:3, 1], M[:3, 2] = r, u, f
    M[:3, 3] = (cam_pos[0], cam_pos[1], -cam_pos[2])
    M[3, :3] = (-np.dot(r, target), -np.dot(u, target), -np.dot(f, target))
    return M


# --------------------------------------------------------------------------------------
# visibility (z-buffered point splat) + hole mask
# --------------------------------------------------------------------------------------
def pack_colour(colour) -> np.ndarray:
    """(N, 3) / (H, W, 3) u8 RGB -> flat uint32 0x00BBGGRR (the low key word of the colour-keyed z-buffers)."""
    c = np.asarray(colour, dtype=np.uint8).reshape(-1, 3).astype(np.uint32)
    return c[:, 0] | (c[:, 1] << 8) | (c[:, 2] << 16)


def splat_ids(u, v, z, out_w: int, out_h: int, near: float = NEAR_PLANE, tie=None) -> np.ndarray:
    """Forward point splat -> id buffer (out_h, out_w) int64, -1 = hole.

    Rule (stereo_rerender.py:746-755,814): target = round-half-even(u, v); keep 0<=x<W,
    0<=y<H; nearest z wins.  The reference's tie order is an unstable argsort (undefined);
    here ties go to the lowest source index, or -- `tie` = one integer per source, the packed
    colour in the product's frame loops -- to the smallest tie value first.  Points with z <= near
    (GL near plane, depth_map_tools.py:1520) or non-finite coordinates are culled before rounding.
    """
    u = np.asarray(u, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        ur = np.rint(u)
        vr = np.rint(v)
        ok = (z > near) & (ur >= 0) & (ur <= out_w - 1) & (vr >= 0) & (vr <= out_h - 1)
    src = np.flatnonzero(ok)
    tgt = vr[src].astype(np.int64) * out_w + ur[src].astype(np.int64)
    if tie is None:
        order = np.lexsort((src, z[src], tgt))  # primary tgt, then z, then source id
    else:
        order = np.lexsort((src, np.asarray(tie).reshape(-1)[src], z[src], tgt))  # tgt, z, tie value, source id
    tgt_s = tgt[order]
    first = np.ones(len(order), dtype=bool)
    first[1:] = tgt_s[1:] != tgt_s[:-1]
    ids = np.full(out_w * out_h, -1, dtype=np.int64)
    ids[tgt_s[first]] = src[order][first]
    return ids.reshape(out_h, out_w)


def resolve(ids: np.ndarray, colour: np.ndarray, bg_rgb=(0, 0, 0), hole_fill=(0, 0, 0), bg_collide=True):
    """id buffer + source colours -> (image u8, hole mask u8 {0,255}).

    Hole = nothing rendered, or (reference quirk) a rendered colour exactly equal to the
    background colour: `bg_mask = all(img == bg_color)` stereo_rerender.py:740,854.  Holes
    are painted `hole_fill`: black in the stereo path (:793), the background itself in
    3d_view_depthfile.py:254.  Colours survive the /255 -> *255 float round trip exactly
    (depth_map_tools.py:1227-1228, stereo_rerender.py:819).
    """
    flat = ids.reshape(-1)
    cols = colour.reshape(-1, 3)
    img = cols[np.maximum(flat, 0)].copy()
    hole = flat < 0
    if bg_collide:
        hole = hole | np.all(img == np.asarray(bg_rgb, dtype=np.uint8), axis=-1)
    img[hole] = np.asarray(hole_fill, dtype=np.uint8)
    h, w = ids.shape
    return img.reshape(h, w, 3), (hole.astype(np.uint8) * 255).reshape(h, w)


def mask_to_rgb(mask_u8: np.ndarray, bg_rgb=(0, 255, 0)) -> np.ndarray:
    """--green_and_black_infill_mask image: bg colour at holes, black elsewhere
    (stereo_rerender.py:787-793,819)."""
    out = np.zeros(mask_u8.shape + (3,), dtype=np.uint8)
    out[mask_u8 != 0] = np.asarray(bg_rgb, dtype=np.uint8)
    return out


def zbuffer_depth(ids: np.ndarray, z: np.ndarray) -> np.ndarray:
    """Rendered depth plane: z of the winner, 0 where nothing was drawn (float32)."""
    flat = ids.reshape(-1)
    out = np.where(flat >= 0, np.asarray(z, dtype=np.float64)[np.maximum(flat, 0)], 0.0)
    return out.astype(np.float32).reshape(ids.shape)


# --------------------------------------------------------------------------------------
# whole-frame drivers
# --------------------------------------------------------------------------------------
def view_uvz(depth_rgb, max_depth, K, M, depth_scale=None, K_out=None, bit16=True, variant="D1",
             of_by_one=False):
    """decode -> (scale) -> unproject -> pose M (4x4) -> project with K_out.  Returns u, v, z'."""
    depth = decode_rgb_depth_frame(depth_rgb, max_depth, bit16, variant)
    if depth_scale is not None:
        depth = apply_depth_scale(depth, depth_scale)
    pts = apply_pose(unproject(depth, K, of_by_one), M)
    return project(pts, K if K_out is None else K_out)


def render_view(depth_rgb, colour, max_depth, K, M, depth_scale=None, K_out=None, out_size=None,
                bg_rgb=(0, 0, 0), hole_fill=(0, 0, 0), tie_colour=False, **kw):
    """One novel view: returns (image, mask, ids).  tie_colour: equal z -> smallest packed colour (frame loops)."""
    h, w = depth_rgb.shape[:2]
    ow, oh = (w, h) if out_size is None else out_size
    u, v, z = view_uvz(depth_rgb, max_depth, K, M, depth_scale, K_out, **kw)
    ids = splat_ids(u, v, z, ow, oh, tie=pack_colour(colour) if tie_colour else None)
    img, mask = resolve(ids, colour, bg_rgb, hole_fill)
    return img, mask, ids


def stereo_frame(depth_rgb, colour, xfov, yfov=None, max_depth=100, pupillary_distance_mm=63,
                 master_xfov=45.0, convergence_depth=None, transform=None, infill_mask=True, tie_colour=False):
    """One iteration of the stereo_rerender.py frame loop (:489-941) in point-splat form.

    Returns (sbs image (H,2W,3) u8, sbs hole mask (H,2W) u8, (ids_left, ids_right)).
    tie_colour: candidates with equal z' are ordered by packed colour instead of source index (the rule of the
    product's colour-keyed frame loops; the reference leaves ties undefined).
    """
    h, w = depth_rgb.shape[:2]
    K = camera_matrix(xfov, yfov, w, h)
    scale = master_fov_depth_scale(master_xfov, xfov)
    ipd = pupillary_distance_mm / 1000
    theta = None
    if convergence_depth is not None and float(convergence_depth) != 0:
        theta = convergence_angle(float(convergence_depth) * scale, ipd)  # :708-718
    T = np.eye(4) if transform is None else np.asarray(transform, dtype=np.float64)
    bg = (0, 255, 0) if infill_mask else (0, 0, 0)
    imgs, masks, idl = [], [], []
    for eye in ("left", "right"):
        M = eye_pose(eye, ipd, theta) @ T
        img, mask, ids = render_view(depth_rgb, colour, max_depth, K, M, depth_scale=scale, bg_rgb=bg, tie_colour=tie_colour)
        imgs.append(img)
        masks.append(mask)
        idl.append(ids)
    return np.concatenate(imgs, axis=1), np.concatenate(masks, axis=1), tuple(idl)


def stereo_frame_uvz(depth_rgb, xfov, yfov=None, max_depth=100, pupillary_distance_mm=63, master_xfov=45.0, convergence_depth=None,
                     transform=None):
    """(u', v', z') of every source pixel in the left and in the right eye, exactly as stereo_frame projects them (the tests use
    them to tell which target pixels a float32 evaluation may legitimately paint differently: rounding boundaries, z ties)."""
    h, w = depth_rgb.shape[:2]
    K = camera_matrix(xfov, yfov, w, h)
    scale = master_fov_depth_scale(master_xfov, xfov)
    ipd = pupillary_distance_mm / 1000
    theta = None
    if convergence_depth is not None and float(convergence_depth) != 0:
        theta = convergence_angle(float(convergence_depth) * scale, ipd)
    T = np.eye(4) if transform is None else np.asarray(transform, dtype=np.float64)
    return [view_uvz(depth_rgb, max_depth, K, eye_pose(eye, ipd, theta) @ T, scale) for eye in ("left", "right")]


def novel_view_frame(depth_rgb, colour, xfov, yfov=None, max_depth=100, cam_pos=(2.0, 2.0, -4.0),
                     target=None, transform=None, center_of_by_one=False, tie_colour=False, want_uvz=False):
    """One iteration of `3d_view_depthfile.py --render` (:133-255) in point-splat form:
    target defaults to the vertex mean (:231; the mesh's vertices sit on the stretched grid
    unless --render_as_pointcloud, :178-182 -> `center_of_by_one`), white background (:254).
    render() scales world Y by fy/fx and projects with fx on both axes
    (depth_map_tools.py:1528-1552)."""
    h, w = depth_rgb.shape[:2]
    K = camera_matrix(xfov, yfov, w, h)
    depth = decode_rgb_depth_frame(depth_rgb, max_depth, True)
    pts = unproject(depth, K, of_by_one=False)
    if transform is not None:
        pts = apply_pose(pts, transform)
    look = pts.mean(axis=0)
    if center_of_by_one:
        grid = unproject(depth, K, of_by_one=True)
        look = (grid if transform is None else apply_pose(grid, transform)).mean(axis=0)
    if target is not None:
        for a in range(3):
            if target[a] is not None:
                look[a] = target[a]
    ext = look_at_extrinsic(np.asarray(cam_pos, dtype=np.float32), look)
    pts = pts * np.array([1.0, K[1][1] / K[0][0], 1.0])
    pts = apply_pose(pts, ext)
    K_r = np.array([[K[0][0], 0, K[0][2]], [0, K[0][0], K[1][2]], [0, 0, 1.0]])
    u, v, z = project(pts, K_r)
    ids = splat_ids(u, v, z, w, h, tie=pack_colour(colour) if tie_colour else None)
    img, mask = resolve(ids, colour, bg_rgb=(255, 255, 255), hole_fill=(255, 255, 255))
    if want_uvz:
        return img, mask, ids, ext, (u, v, z)
    return img, mask, ids, ext


# --------------------------------------------------------------------------------------
# convergence list preparation (whole clip, before any sharding)
# --------------------------------------------------------------------------------------
def fill_nan_with_closest(values):
    """stereo_rerender.py:243-250: each NaN takes the value of the nearest non-NaN index
    (ties -> the earlier index)."""
    vals = list(values)
    good = [i for i, x in enumerate(vals) if not math.isnan(x)]
    if good:
        for i, x in enumerate(vals):
            if math.isnan(x):
                vals[i] = vals[min(good, key=lambda j: abs(j - i))]
    return vals


def smooth_convergence(values):
    """stereo_rerender.py:252-268: Savitzky-Golay (order 2, window <= 99, odd) over the list
    extended by its own last <= 50 samples; the extension is cut off again."""
    from scipy.signal import savgol_filter

    y = np.array(values)
    n_tail = min(50, len(y))
    y_ext = np.concatenate([y, y[-n_tail:]])
    win = min(100, len(y_ext))
    if win % 2 == 0:
        win -= 1
    sm = savgol_filter(y_ext, window_length=win, polyorder=2)
    return sm[:-n_tail] if n_tail > 0 else sm


def convergence_depth_of_frame(depth_rgb, max_depth=100, mask=None):
    """find_convergence_depth.py:56-80: mean of the D3-decoded depth (under mask > 240)."""
    d = decode_rgb_depth_frame(depth_rgb, max_depth, True, "D3")
    if mask is not None:
        d = d[mask > 240]
    return float(d.mean()) if d.size else float("nan")


# --------------------------------------------------------------------------------------
# PLY export (config 1)
# --------------------------------------------------------------------------------------
def ply_points_of_frame(depth_rgb, colour, xfov, yfov=None, max_depth=100, transform=None):
    """convert_metric_depth_video_to_other_format.py:646-652,692-695,743-749: D2 decode,
    of_by_one=True (the get_mesh_from_depth_map default, depth_map_tools.py:1105), optional
    pose; colours /255 (depth_map_tools.py:1227-1228).  Returns (xyz f64 (N,3), rgb u8 (N,3))."""
    h, w = depth_rgb.shape[:2]
    K = camera_matrix(xfov, yfov, w, h)
    depth = decode_rgb_depth_frame(depth_rgb, max_depth, True, "D2")
    pts = unproject(depth, K, of_by_one=True)
    if transform is not None:
        pts = apply_pose(pts, transform)
    return pts, colour.reshape(-1, 3).copy()


def write_ply(path, xyz: np.ndarray, rgb_u8: np.ndarray):
    """Binary little-endian PLY, double x y z + uchar red green blue: the layout of Open3D's
    legacy writer behind o3d.io.write_point_cloud (convert_...py:748-749).  Open3D is absent
    here, so the byte layout is restated from its published format, not pinned."""
    n = len(xyz)
    rec = np.empty(n, dtype=[("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("r", "u1"), ("g", "u1"), ("b", "u1")])
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rec["r"], rec["g"], rec["b"] = rgb_u8[:, 0], rgb_u8[:, 1], rgb_u8[:, 2]
    header = (
        "ply\nformat binary_little_endian 1.0\ncomment Created by Open3D\n"
        f"element vertex {n}\nproperty double x\nproperty double y\nproperty double z\n"
        "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n"
    )
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(rec.tobytes())


def read_ply(path):
    with open(path, "rb") as fh:
        blob = fh.read()
    end = blob.index(b"end_header\n") + len(b"end_header\n")
    n = int([ln for ln in blob[:end].decode("ascii").splitlines() if ln.startswith("element vertex")][0].split()[-1])
    rec = np.frombuffer(blob[end:], dtype=[("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("r", "u1"), ("g", "u1"), ("b", "u1")], count=n)
    return np.stack([rec["x"], rec["y"], rec["z"]], axis=1), np.stack([rec["r"], rec["g"], rec["b"]], axis=1)
