"""Host-side scalar geometry of the path (per frame, a handful of float64 values): camera matrix,
eye poses, convergence handling, look-at.  These are the reference's own small NumPy helpers
(cited per function); nothing here touches pixels."""
from __future__ import annotations

import math

import numpy as np

NEAR_PLANE = 1e-4  # depth_map_tools.py:1520


def compute_camera_matrix(fov_horizontal_deg, fov_vertical_deg, image_width, image_height) -> np.ndarray:
    """depth_map_tools.py:902-934.  One FOV may be None: that axis copies the other focal length."""
    if fov_horizontal_deg is None and fov_vertical_deg is None:
        raise ValueError("at least one of the horizontal / vertical field of view is required")
    focal = {}
    for axis, fov, size in (("x", fov_horizontal_deg, image_width), ("y", fov_vertical_deg, image_height)):
        if fov is not None:
            focal[axis] = size / (2 * np.tan(np.deg2rad(fov) / 2))
    fx = focal.get("x", focal.get("y"))
    fy = focal.get("y", focal.get("x"))
    return np.array([[fx, 0, image_width / 2], [0, fy, image_height / 2], [0, 0, 1]], dtype=np.float64)


def fov_from_camera_matrix(mat):
    """depth_map_tools.py:1640-1649."""
    width, height = mat[0][2] * 2, mat[1][2] * 2
    return (np.rad2deg(2 * np.arctan2(width, 2 * mat[0][0])), np.rad2deg(2 * np.arctan2(height, 2 * mat[1][1])))


def master_fov_depth_scale(master_xfov_deg: float, xfov_deg: float) -> float:
    """stereo_rerender.py:537-538."""
    return 1.0 / (math.tan(math.radians(master_xfov_deg / 2)) / math.tan(math.radians(xfov_deg / 2)))


def convergence_angle(distance: float, pupillary_distance: float) -> float:
    """stereo_rerender.py:94-112."""
    if distance == 0:
        raise ValueError("Distance must be non-zero to compute a valid angle.")
    return math.atan((pupillary_distance / 2) / distance)


def rotation_about_y(angle: float) -> np.ndarray:
    """Open3D get_rotation_matrix_from_xyz((0, angle, 0)) as used at stereo_rerender.py:719-720."""
    c, s = math.cos(angle), math.sin(angle)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def stereo_eye_pose(eye: str, ipd_m: float, conv_angle=None) -> np.ndarray:
    """Frame space -> eye camera.  left: Ry(-theta) then +ipd/2 (stereo_rerender.py:723-725);
    right: Ry(+theta) then -ipd/2 (:831-836)."""
    sign = 1.0 if eye == "left" else -1.0
    pose = np.eye(4)
    if conv_angle:
        pose[:3, :3] = rotation_about_y(-sign * conv_angle)
    pose[0, 3] = sign * ipd_m / 2
    return pose


def cam_look_at(cam_pos, target, up=np.array([0.0, 1.0, 0.0])) -> np.ndarray:
    """depth_map_tools.py:1618-1638 (the reference's own convention: r, u, f as columns of the upper
    3x3, translation (px, py, -pz)); Open3D reads the upper 3x4 as world -> camera."""
    cam_pos = np.asarray(cam_pos)
    fwd = np.asarray(target, dtype=np.float64) - cam_pos
    fwd = fwd / np.linalg.norm(fwd)
    right = np.cross(up, fwd)
    right = right / np.linalg.norm(right)
    true_up = np.cross(fwd, right)
    target = np.asarray(target, dtype=np.float64)
    return np.array([
        [right[0], true_up[0], fwd[0], cam_pos[0]],
        [right[1], true_up[1], fwd[1], cam_pos[1]],
        [right[2], true_up[2], fwd[2], -cam_pos[2]],
        [-np.dot(right, target), -np.dot(true_up, target), -np.dot(fwd, target), 1.0],
    ], dtype=float)


def gl_look_at(eye, target, up) -> np.ndarray:
    """depth_map_tools.py:1599-1616: OpenGL view matrix (float32) looking from `eye` at `target`."""
    f = np.asarray(target) - np.asarray(eye)
    f = f / np.linalg.norm(f)
    s = np.cross(f, up)
    s = s / np.linalg.norm(s)
    u = np.cross(s, f)
    M = np.eye(4, dtype=np.float32)
    M[0, :3], M[1, :3], M[2, :3] = s, u, -f
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = -np.asarray(eye)
    return M @ T


def get_cam_view(side_offset, convergence_angle_rad=0.0, reverse=False) -> np.ndarray:
    """depth_map_tools.py:226-245: eye view matrix = Ry(angle) @ T(offset) @ look-down-(-z) (or its reverse order)."""
    def rot_y(a):
        c, s = np.cos(a), np.sin(a)
        return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float32)

    def shift_x(x):
        T = np.eye(4, dtype=np.float32)
        T[:3, 3] = [x, 0, 0]
        return T

    eye = np.array([0, 0, 0], dtype=np.float32)
    base = gl_look_at(eye, eye + np.array([0, 0, -1], dtype=np.float32), np.array([0, 1, 0], dtype=np.float32))
    if not reverse:
        return rot_y(convergence_angle_rad) @ shift_x(side_offset) @ base
    return shift_x(-side_offset) @ rot_y(-convergence_angle_rad) @ base


def open_cv_w2c_to_gl_view(transform_to_ref) -> np.ndarray:
    """depth_map_tools.py:62-75: OpenCV camera-to-reference pose -> OpenGL view matrix (y and z axes flipped)."""
    w2c = np.linalg.inv(transform_to_ref)
    A = np.diag([1, -1, -1, 1]).astype(np.float32)
    return np.linalg.inv(A @ w2c @ A)


def rebase_transformations(transformations, lock_frame: int):
    """stereo_rerender.py:369-373: T_i <- T_i @ inv(T_lock) when a lock frame other than 0 is given."""
    mats = [np.asarray(t, dtype=np.float64) for t in transformations]
    if lock_frame != 0:
        inv = np.linalg.inv(mats[lock_frame])
        mats = [m @ inv for m in mats]
    return mats


def fill_nan_with_closest(values):
    """stereo_rerender.py:243-250."""
    vals = list(values)
    known = [i for i, v in enumerate(vals) if not math.isnan(v)]
    if known:
        for i, v in enumerate(vals):
            if math.isnan(v):
                vals[i] = vals[min(known, key=lambda k: abs(k - i))]
    return vals


def curve_fit(values):
    """stereo_rerender.py:252-268: Savitzky-Golay smoothing of the whole-clip convergence list; must run
    before the clip is sharded across GPUs because the filter window spans frames."""
    from scipy.signal import savgol_filter

    y = np.array(values)
    n_tail = min(50, len(y))
    y_ext = np.concatenate([y, y[-n_tail:]])
    window = min(100, len(y_ext))
    window -= 1 - window % 2
    smooth = savgol_filter(y_ext, window_length=window, polyorder=2)
    smooth = smooth[:-n_tail] if n_tail > 0 else smooth
    assert len(smooth) == len(y), f"curve_fit output length {len(smooth)} != input length {len(y)}"
    return smooth
