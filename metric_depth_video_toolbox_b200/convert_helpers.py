"""The module-level helpers of convert_metric_depth_video_to_other_format.py (:14-470), importable under the reference's
names.  Two are plain NumPy formulas and are implemented (pinned to the reference's outputs, tests/golden/convert_helpers.npz);
the rest belong to the script's sparse triangulation / Blender-Alembic export, which is outside the dense per-frame path
(DESIGN.md 8): they refuse with that reason."""
from __future__ import annotations

import numpy as np


def float_image_to_byte_image(float_image, max_value=10.0, scale=255, log_scale=5):
    """Log-scaled 8-bit preview of a float image (convert_...py:14-29): clip to [1e-4, max_value], log(v * log_scale + 1)
    normalised by its value at max_value, times `scale`, clipped and truncated to uint8."""
    v = np.clip(float_image, 0.0001, max_value)
    coded = np.log(v * log_scale + 1) / np.log(max_value * log_scale + 1) * scale
    return np.clip(coded, 0, scale).astype(np.uint8)


def estimate_scale_shift(depth, depth_target):
    """Least-squares (scale, shift) of 1 / depth_target = scale * (1 / depth) + shift over the pixels where both depths are
    positive (convert_...py:443-470)."""
    ok = (depth > 0) & (depth_target > 0)
    x, y = 1.0 / depth[ok], 1.0 / depth_target[ok]
    (scale, shift), _, _, _ = np.linalg.lstsq(np.vstack([x, np.ones_like(x)]).T, y, rcond=None)
    return scale, shift


class OutOfScope(NotImplementedError):
    """A helper of the export script's sparse triangulation / Alembic branch."""


def _refuse(name: str, what: str):
    raise OutOfScope(f"convert_metric_depth_video_to_other_format.{name}: {what} -- outside the dense per-frame path this package "
                     f"implements (SURVEY.md section 8)")


def compute_weights_chunked(directions, chunk_size=1024):
    _refuse("compute_weights_chunked", "ray-pair weights of the sparse triangulation (--triangulate)")


def best_intersection_point_vectorized_weighted(points, directions, weights=None):
    _refuse("best_intersection_point_vectorized_weighted", "least-squares ray intersection of the sparse triangulation")


def find_nearby_points(points_3d, i, threshold=0.01, exclude_self=True):
    _refuse("find_nearby_points", "neighbour search of the sparse triangulation")


def merge_global_points(global_3d_points, remaped_points):
    _refuse("merge_global_points", "union-find merge of tracked points")


def add_open3d_mesh(o3d_mesh, object_name="ImportedMesh"):
    _refuse("add_open3d_mesh", "Blender (bpy) scene construction")


def add_point_cloud(point_cloud, point_colors=None, object_name="PointCloud"):
    _refuse("add_point_cloud", "Blender (bpy) scene construction")


def assign_vertex_color_material(obj, vcol_name="Col"):
    _refuse("assign_vertex_color_material", "Blender (bpy) materials")


def create_camera_alembic(transforms, output_file, fps=24.0, camera_name="TrackedCamera", intrinsic_matrix=None, resolution=(1920, 1080),
                          point_cloud_points=None, point_cloud_colors=None, point_cloud_name="PointCloud", open3d_mesh=None,
                          open3d_mesh_name="ImportedMesh", blend_filepath=None):
    _refuse("create_camera_alembic", "Alembic camera export through Blender (bpy)")
