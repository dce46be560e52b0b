"""Launcher with the reference script's name and command line (see metric_depth_video_toolbox_b200/cli/convert_format.py); the
script's module-level helpers are importable under their names (convert_helpers.py: two NumPy formulas implemented, the
sparse triangulation / Blender-Alembic ones refuse with the reason)."""
from metric_depth_video_toolbox_b200.cli.convert_format import main
from metric_depth_video_toolbox_b200.convert_helpers import (  # noqa: F401
    add_open3d_mesh, add_point_cloud, assign_vertex_color_material, best_intersection_point_vectorized_weighted, compute_weights_chunked,
    create_camera_alembic, estimate_scale_shift, find_nearby_points, float_image_to_byte_image, merge_global_points)

if __name__ == "__main__":
    raise SystemExit(main())
