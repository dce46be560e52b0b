"""Drop-in for the hot-path surface of the reference's `depth_map_tools` module: same function names,
argument order and results for camera matrices, unprojection, point transforms / projection, the
"mesh" built from a depth map and `render()` into a virtual camera.

Two things differ by design:

  * The geometry objects are light device-resident handles (`DepthMesh`, `PointCloud`) instead of Open3D
    objects: a `DepthMesh` is the depth plane + colours + intrinsics + an accumulated 4x4 pose, and its
    `transform / rotate / translate` only update that matrix.  Vertices are materialised (one kernel) only
    when `.vertices` is read.
  * `render()` is a z-buffered forward point splat on the GPU (nearest z wins, round-half-even pixel
    snapping, near plane 1e-4), the visibility rule of the reference's own point painter
    (stereo_rerender.py:746-755,814), instead of Open3D's OpenGL rasteriser -- which exists only on Windows
    (depth_map_tools.py:1461).

Array arguments may be NumPy arrays or CUDA tensors; NumPy in -> NumPy out.  No CPU implementation.
"""
from __future__ import annotations

import contextlib
import math
import time
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import geometry as geo
from . import ops
from .depth_frames_helper import _device, _down, _up
from .geometry import compute_camera_matrix, fov_from_camera_matrix  # noqa: F401  (depth_map_tools.py:902-934,1640-1649)

NEAR_PLANE = geo.NEAR_PLANE
zero_identity_matrix = np.identity(4)  # depth_map_tools.py:1185
# module globals of the reference's render() (depth_map_tools.py:1417-1421): kept so that `depth_map_tools.vis` etc. resolve;
# the splat renderer holds no window / visualiser state
vis = None
v_h = None
v_w = None
rend = None
use_ofscreen = True


@contextlib.contextmanager
def timer(name='not named'):
    """depth_map_tools.py:13-18."""
    start = time.perf_counter()
    yield
    print(f"{name}: {time.perf_counter() - start:.6f} seconds")


# ---------------------------------------------------------------------------------------------
# small matrix helpers (host scalars)
# ---------------------------------------------------------------------------------------------
def rotation_y(angle_rad):
    """depth_map_tools.py:209-219 (4x4 float32)."""
    c, s = np.cos(angle_rad), np.sin(angle_rad)
    return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float32)


def translation_matrix(x, y, z):
    """depth_map_tools.py:221-224."""
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [x, y, z]
    return T


def convergence_angle(distance, pupillary_distance):
    """depth_map_tools.py:247-262."""
    return geo.convergence_angle(distance, pupillary_distance)


def cam_look_at(cam_pos, target, up=np.array([0.0, 1.0, 0.0])):
    """depth_map_tools.py:1618-1638."""
    return geo.cam_look_at(cam_pos, target, up)


def gl_look_at(eye, target, up):
    """depth_map_tools.py:1599-1616."""
    return geo.gl_look_at(eye, target, up)


def get_cam_view(side_offset, convergence_angle_rad=0.0, reverse=False):
    """depth_map_tools.py:226-245."""
    return geo.get_cam_view(side_offset, convergence_angle_rad, reverse)


def open_cv_w2c_to_gl_view(transform_to_ref):
    """depth_map_tools.py:62-75."""
    return geo.open_cv_w2c_to_gl_view(transform_to_ref)


def reject_outliers(data, m=1):
    """depth_map_tools.py:1037-1038."""
    return abs(data - np.mean(data)) < m * np.std(data)


def apply_side_view_to_paralax_mask(parallax_mask, normals, right):
    """depth_map_tools.py:195-207: keep the mask where the normal's x component looks towards (right) / away from the side."""
    right_dot = normals[..., 0]
    cos_threshold = np.cos(np.deg2rad(90.0))
    return parallax_mask & ((right_dot > cos_threshold) if right else (right_dot < cos_threshold))


def get_rotation_matrix_from_xyz(angles):
    """Open3D Geometry3D.get_rotation_matrix_from_xyz: R = Rx(a) @ Ry(b) @ Rz(c)."""
    a, b, c = (float(v) for v in angles)
    rx = np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
    ry = np.array([[math.cos(b), 0, math.sin(b)], [0, 1, 0], [-math.sin(b), 0, math.cos(b)]])
    rz = np.array([[math.cos(c), -math.sin(c), 0], [math.sin(c), math.cos(c), 0], [0, 0, 1]])
    return rx @ ry @ rz


# ---------------------------------------------------------------------------------------------
# geometry handles
# ---------------------------------------------------------------------------------------------
class _Posed:
    """Accumulated rigid/affine pose; Open3D's transform / rotate / translate semantics."""

    def __init__(self):
        self.pose = np.eye(4)

    def transform(self, T):
        self.pose = np.asarray(T, dtype=np.float64) @ self.pose
        return self

    def rotate(self, R, center=(0.0, 0.0, 0.0)):
        c = np.asarray(center, dtype=np.float64)
        T = np.eye(4)
        T[:3, :3] = np.asarray(R, dtype=np.float64)
        T[:3, 3] = c - T[:3, :3] @ c
        return self.transform(T)

    def translate(self, t, relative=True):
        if not relative:
            t = np.asarray(t, dtype=np.float64) - self.get_center()
        T = np.eye(4)
        T[:3, 3] = np.asarray(t, dtype=np.float64)
        return self.transform(T)

    def _pose_or_none(self):
        return None if np.array_equal(self.pose, np.eye(4)) else self.pose

    get_rotation_matrix_from_xyz = staticmethod(get_rotation_matrix_from_xyz)


class DepthMesh(_Posed):
    """What get_mesh_from_depth_map returns: the grid "mesh" of one depth frame, device resident.

    depth: (H, W) float32 CUDA; colour: (H, W, 3) u8 CUDA or None; K: float64 3x3; of_by_one as given to
    create_point_cloud_from_depth (affects `.vertices` / `get_center()` only -- render() splats the exact
    pixel grid, because the stretch compensates for triangle rasterisation which the splat replaces)."""

    def __init__(self, depth: torch.Tensor, colour: Optional[torch.Tensor], K: np.ndarray, of_by_one: bool):
        super().__init__()
        self.depth, self.colour, self.K, self.of_by_one = depth, colour, np.asarray(K, dtype=np.float64), bool(of_by_one)
        self.removed = None  # optional (N,) bool CUDA tensor: vertices taken out by convert_mesh_to_pcd

    @property
    def height(self):
        return self.depth.shape[0]

    @property
    def width(self):
        return self.depth.shape[1]

    def __len__(self):
        return self.depth.numel()

    def source(self, of_by_one: Optional[bool] = None):
        return ops.make_source(self.width, self.height, self.K, decoder="F32", of_by_one=self.of_by_one if of_by_one is None else of_by_one)

    def vertices_device(self) -> torch.Tensor:
        return ops.unproject(self.depth, self.source(), self._pose_or_none(), torch.float64, self.K)

    @property
    def vertices(self) -> np.ndarray:
        """(N, 3) float64, bit-identical to create_point_cloud_from_depth (+ pose)."""
        return self.vertices_device().cpu().numpy()

    @property
    def vertex_colors(self) -> Optional[np.ndarray]:
        """(N, 3) float64 in [0, 1] (depth_map_tools.py:1227-1228)."""
        return None if self.colour is None else self.colour.reshape(-1, 3).cpu().numpy() / 255.0

    def get_center(self) -> np.ndarray:
        """Mean of the vertices (Open3D get_center) from one fused decode-free reduction kernel."""
        s = ops.centroid_sums(self.depth, self.source(), self.K, self._pose_or_none()).cpu().numpy()
        return s[:3] / s[3]


class PointCloud(_Posed):
    """pts_2_pcd result: explicit points (N, 3) float64 and optional colours (N, 3) in [0, 1]."""

    def __init__(self, points, colors=None, normals=None):
        super().__init__()
        self._points, _ = _up(np.asarray(points).reshape(-1, 3) if not isinstance(points, torch.Tensor) else points, torch.float64)
        self._colors = None if colors is None else _up(np.asarray(colors).reshape(-1, 3) if not isinstance(colors, torch.Tensor) else colors,
                                                       torch.float64)[0]
        self._normals = None if normals is None else np.asarray(normals)

    def __len__(self):
        return self._points.shape[0]

    def points_device(self) -> torch.Tensor:
        pose = self._pose_or_none()
        return self._points if pose is None else ops.transform_points(self._points, pose)

    @property
    def points(self) -> np.ndarray:
        return self.points_device().cpu().numpy()

    @property
    def colors(self) -> Optional[np.ndarray]:
        return None if self._colors is None else self._colors.cpu().numpy()

    @property
    def normals(self):
        return self._normals

    def colours_u8_device(self) -> torch.Tensor:
        if self._colors is None:
            return torch.zeros((len(self), 3), dtype=torch.uint8, device=self._points.device)
        return (self._colors * 255).to(torch.uint8)  # the reference's (x * 255).astype(uint8): truncation

    def get_center(self) -> np.ndarray:
        return self.points_device().mean(dim=0).cpu().numpy() if len(self) else np.zeros(3)


class PointMesh(PointCloud):
    """create_mesh_from_point_cloud result: explicit grid-organised vertices (`grid` = (H, W)); drawn by render() as
    its vertices (the splat replaces triangle rasterisation)."""
    grid = None

    @property
    def vertices(self) -> np.ndarray:
        return self.points

    @property
    def vertex_colors(self):
        return self.colors


# ---------------------------------------------------------------------------------------------
# module functions (signatures as in the reference)
# ---------------------------------------------------------------------------------------------
def create_point_cloud_from_depth(depth_image, intrinsics, of_by_one=False):
    """depth_map_tools.py:1112-1133 -> (points (H*W, 3) float64, height, width)."""
    depth, as_np = _up(depth_image, torch.float32)
    h, w = depth.shape
    K = np.asarray(intrinsics, dtype=np.float64)
    pts = ops.unproject(depth, ops.make_source(w, h, K, decoder="F32", of_by_one=of_by_one), None, torch.float64, K)
    return _down(pts, as_np), h, w


def get_mesh_from_depth_map(depth_map, cam_mat, color_frame=None, inp_mesh=None, remove_edges=False, mask=None,
                            invalid_color=None, of_by_one=True, return_normals_of_removed=False):
    """depth_map_tools.py:1104-1110.  Returns (mesh, used_indices) or, with return_normals_of_removed,
    (mesh, unused_indices, normals_of_removed).  `inp_mesh` is reused (its buffers are overwritten)."""
    if mask is not None or invalid_color is not None:
        raise NotImplementedError("mask / invalid_color belong to the mesh-cell filter, which the point splat does not have")
    depth, _ = _up(depth_map, torch.float32)
    colour = None if color_frame is None else _up(color_frame, torch.uint8)[0]
    if isinstance(inp_mesh, DepthMesh) and inp_mesh.depth.shape == depth.shape:
        mesh = inp_mesh
        mesh.depth, mesh.colour, mesh.K, mesh.of_by_one, mesh.pose, mesh.removed = depth, colour, np.asarray(cam_mat, np.float64), bool(of_by_one), np.eye(4), None
    else:
        mesh = DepthMesh(depth, colour, cam_mat, of_by_one)
    n = depth.numel()
    if remove_edges:
        from . import edges

        unused, normals = edges.edge_vertices(mesh)
        if return_normals_of_removed:
            return mesh, unused, normals
        used = np.ones(n, dtype=bool)
        used[unused] = False
        return mesh, np.where(used)[0]
    if return_normals_of_removed:
        return mesh, np.zeros(0, dtype=np.int64), []
    return mesh, np.arange(n)


def calculate_normals(depth, K):
    """depth_map_tools.py:20-60 -> (H, W, 3) float32 unit normals (y, z flipped), one kernel.  The reference computes in
    the depth array's float32 (its only caller hands it a float32 plane, :282); other dtypes are converted first."""
    d, as_np = _up(depth, torch.float32)
    return _down(ops.calculate_normals(d.contiguous(), np.asarray(K, dtype=np.float64)), as_np)


def create_mesh_from_point_cloud(points, height, width, image_frame=None, inp_mesh=None, remove_edges=False, mask=None,
                                 angle_threshold_deg=89.0, invalid_color=None, background_edge_mask_expandansions=0,
                                 return_normals_of_removed=False):
    """depth_map_tools.py:1186-1416 for grid-organised points (what create_point_cloud_from_depth returns).  The mesh handle
    is a `PointMesh` (vertices + colours + pose: render() splats it like a point cloud); the edge test
    (`remove_edges`, 89 degrees by default) runs on the GPU and the results come back in the reference's shapes:
    (mesh, used_indices) or, with return_normals_of_removed, (mesh, unused_indices, normals_of_removed)."""
    if mask is not None or invalid_color is not None or background_edge_mask_expandansions:
        raise NotImplementedError("mask / invalid_color / background_edge_mask_expandansions belong to the mesh-cell filter, which the point "
                                  "splat does not have")
    pts, _ = _up(np.asarray(points).reshape(-1, 3) if not isinstance(points, torch.Tensor) else points.reshape(-1, 3), torch.float64)
    height, width = int(height), int(width)
    if pts.shape[0] != height * width:
        raise ValueError(f"{pts.shape[0]} points do not form a {height}x{width} grid")
    colors = None if image_frame is None else np.asarray(image_frame).reshape(-1, 3) / 255.0
    if isinstance(inp_mesh, PointMesh) and len(inp_mesh) == pts.shape[0]:
        mesh = inp_mesh
        mesh.__init__(pts, colors)
    else:
        mesh = PointMesh(pts, colors)
    mesh.grid = (height, width)
    n = pts.shape[0]
    if remove_edges:
        flags, normals = ops.edge_vertices_xyz(pts.contiguous(), height, width, True, angle_threshold_deg)
        idx = torch.nonzero(flags.reshape(-1), as_tuple=False).reshape(-1)
        unused = idx.cpu().numpy().astype(np.int64)
        if return_normals_of_removed:
            return mesh, unused, normals.reshape(-1, 3)[idx].cpu().numpy()
        used = np.ones(n, dtype=bool)
        used[unused] = False
        return mesh, np.where(used)[0]
    if return_normals_of_removed:
        return mesh, np.zeros(0, dtype=np.int64), []
    return mesh, np.arange(n)


def transform_points(points, transform):
    """depth_map_tools.py:977-1004."""
    pts, as_np = _up(points, torch.float64)
    return _down(ops.transform_points(pts.reshape(-1, 3), np.asarray(transform, dtype=np.float64)), as_np)


def project_3d_points_to_2d(t3d_points, cam_mat, distCoeffs=np.array([0, 0, 0, 0])):
    """depth_map_tools.py:1057-1060.  Zero distortion only (all the scripts pass)."""
    if np.any(np.asarray(distCoeffs) != 0):
        raise NotImplementedError("lens distortion is not part of the GPU path")
    pts, as_np = _up(t3d_points, torch.float64)
    K32 = np.asarray(cam_mat).astype(np.float32).astype(np.float64)  # the reference hands cv2 a float32 matrix
    return _down(ops.project_points(pts.reshape(-1, 3), K32), as_np).squeeze()


def pts_2_pcd(points, colors=None, ids=None, normals=None):
    """depth_map_tools.py:1040-1055."""
    return PointCloud(points, colors, normals)


def convert_mesh_to_pcd(mesh, points_to_remove, input_pcd):
    """depth_map_tools.py:1086-1102: the mesh's vertices as a point cloud; removed vertices are parked behind
    the camera at (-0.2, -0.2, -0.2) in the reference -- here they are simply flagged and skipped."""
    removed = None
    if points_to_remove is not None and len(points_to_remove):
        removed = torch.zeros(len(mesh), dtype=torch.bool, device=mesh.depth.device)
        removed[torch.as_tensor(np.asarray(points_to_remove), device=mesh.depth.device, dtype=torch.long)] = True
    mesh.removed = removed
    return mesh


# ---------------------------------------------------------------------------------------------
# render
# ---------------------------------------------------------------------------------------------
_zbufs = {}


def _zbuf(w: int, h: int, device) -> torch.Tensor:
    key = (w, h, str(device), torch.cuda.current_stream(device).cuda_stream)
    z = _zbufs.get(key)
    if z is None:
        if len(_zbufs) > 8:
            _zbufs.clear()
        z = _zbufs[key] = ops.new_zbuf(1, w, h, device)
    return z


def render_device(objects: Sequence, cam_mat, w=None, h=None, extrinsic_matric=np.eye(4), bg_color=np.array([0, 0, 0]),
                  want_depth: bool = True):
    """render() without the host round trip: (rgb u8 (h, w, 3), hole mask u8 (h, w), depth f32 (h, w) or None),
    CUDA tensors.  `bg_color` in [0, 1] like the reference; holes are painted with it."""
    K = np.asarray(cam_mat, dtype=np.float64)
    if w is None:
        w, h = K[0][2] * 2, K[1][2] * 2
    w, h = int(w), int(h)
    dev = _device()
    ext = np.asarray(extrinsic_matric, dtype=np.float64)
    zbuf = _zbuf(w, h, dev)
    colour_tables: List[torch.Tensor] = []
    offset = 0
    # The reference works around Open3D ignoring fy by scaling the geometry's Y by fy/fx *before* the extrinsic and
    # projecting with fx on both axes (depth_map_tools.py:1528-1552); identical to a plain pinhole when the
    # extrinsic does not mix Y with X/Z, kept verbatim otherwise.  Open3D reads the upper 3x4 of the extrinsic.
    y_scale = np.diag([1.0, K[1, 1] / K[0, 0], 1.0, 1.0])
    for obj in objects:
        view = ops.ViewSpec(ext[:3, :4] @ y_scale @ obj.pose, K[0, 0], K[0, 0], K[0, 2], K[1, 2])
        if isinstance(obj, DepthMesh):
            depth = obj.depth if obj.removed is None else obj.depth.masked_fill(obj.removed.view_as(obj.depth), 0.0)
            ops.project_splat(depth, obj.source(of_by_one=False), [view], w, h, zbuf, NEAR_PLANE, id_offset=offset)
            colour_tables.append(obj.colour.reshape(-1, 3) if obj.colour is not None
                                 else torch.full((len(obj), 3), 255, dtype=torch.uint8, device=dev))
        elif isinstance(obj, PointCloud):
            ops.splat_points(obj._points.to(torch.float32), [view], w, h, zbuf, NEAR_PLANE, id_offset=offset)
            colour_tables.append(obj.colours_u8_device())
        else:
            raise TypeError(f"render() takes DepthMesh / PointCloud objects, got {type(obj).__name__}")
        offset += len(obj)
    table = colour_tables[0] if len(colour_tables) == 1 else torch.cat(colour_tables)
    bg = tuple(int(v) for v in (np.asarray(bg_color, dtype=np.float64) * 255).astype(np.uint8))
    rgb, mask, depth, _ = ops.resolve(zbuf[0], table.contiguous(), bg, bg, ops.FLAG_RESET_ZBUF, want_depth=want_depth)
    return rgb, mask, depth


def render(objects, cam_mat, depth=False, w=None, h=None, extrinsic_matric=np.eye(4), bg_color=np.array([0, 0, 0])):
    """depth_map_tools.py:1422-1597.  depth=False -> float32 RGB (h, w, 3) in [0, 1]; depth=True -> float32 depth
    (0 where nothing was drawn); depth=-2 -> (rgb, depth)."""
    rgb, _, z = render_device(objects, cam_mat, w, h, extrinsic_matric, bg_color, want_depth=depth is not False)
    if depth is True:
        return z.cpu().numpy()
    image = (rgb.to(torch.float32) / 255.0).cpu().numpy()
    if depth == -2:
        return image, z.cpu().numpy()
    return image


# ---------------------------------------------------------------------------------------------
# names of the reference module that are outside the dense per-frame path: importable, refuse with the reason
# ---------------------------------------------------------------------------------------------
def _outside_the_path(name: str, where: str, why: str):
    def refuse(*_args, **_kwargs):
        raise NotImplementedError(f"depth_map_tools.{name} ({where}) is not part of the GPU per-frame path: {why}")

    refuse.__name__ = name
    refuse.__doc__ = f"{where}: {why} (refuses when called)."
    return refuse


_GL = "the reference's own OpenGL rasteriser pipeline, unused by stereo_rerender / 3d_view_depthfile / convert_...; render() is the splat"
_SPARSE = "sparse host-side tracking / registration maths on a few hundred points (SciPy / OpenCV solvers), not per-pixel work"
_O3D = "an Open3D-only utility (voxel down-sampling / interactive window)"
mesh_from_depth_and_rgb = _outside_the_path("mesh_from_depth_and_rgb", "depth_map_tools.py:265-466", _GL)
mesh_maker_helper_make_corner_unclamped = _outside_the_path("mesh_maker_helper_make_corner_unclamped", "depth_map_tools.py:468-470", _GL)
mesh_maker_helper_make_corner_with_mask = _outside_the_path("mesh_maker_helper_make_corner_with_mask", "depth_map_tools.py:472-485", _GL)
remap_ids_to_img = _outside_the_path("remap_ids_to_img", "depth_map_tools.py:487-539", _GL)
steep_disparity_lr = _outside_the_path("steep_disparity_lr", "depth_map_tools.py:541-571", _GL)
steep_mask_disparity = _outside_the_path("steep_mask_disparity", "depth_map_tools.py:573-609", _GL)
generate_normal_bg_image = _outside_the_path("generate_normal_bg_image", "depth_map_tools.py:611-656", _GL)
gl_render = _outside_the_path("gl_render", "depth_map_tools.py:660-865", _GL)
open_gl_projection_from_camera_matrix = _outside_the_path("open_gl_projection_from_camera_matrix", "depth_map_tools.py:867-900", _GL)
frustum_planes = _outside_the_path("frustum_planes", "depth_map_tools.py:82-134", _SPARSE)
frusta_intersect = _outside_the_path("frusta_intersect", "depth_map_tools.py:136-193", _SPARSE)
svd = _outside_the_path("svd", "depth_map_tools.py:937-975", _SPARSE)
pnpSolve_ransac = _outside_the_path("pnpSolve_ransac", "depth_map_tools.py:1006-1035", _SPARSE)
project_2d_points_to_3d = _outside_the_path("project_2d_points_to_3d", "depth_map_tools.py:1062-1084", _SPARSE)
perspective_aware_down_sample = _outside_the_path("perspective_aware_down_sample", "depth_map_tools.py:1136-1183", _O3D)
draw = _outside_the_path("draw", "depth_map_tools.py:1652-1658", _O3D)
