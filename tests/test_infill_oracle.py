"""CPU: the infill-mask oracle (oracle/infill_oracle.py) against golden vectors produced by RUNNING the reference
(oracle/make_infill_golden.py: depth_map_tools.create_mesh_from_point_cloud through a fake Open3D shim and the
stereo_rerender.py mask lines by line range) -- bit for bit, including OpenCV's TELEA inpainting and the masked blur."""
import os

import numpy as np
import pytest

from oracle import infill_oracle as io
from oracle import mdvt_oracle as orc


@pytest.mark.parametrize("tag", ["plain", "posed"])
def test_infill_oracle_equals_reference_outputs(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "infill_mask.npz"))
    w, h, xfov, conv = g[tag + "_params"]
    w, h = int(w), int(h)
    K = orc.camera_matrix(xfov, None, w, h)
    scale = orc.master_fov_depth_scale(45.0, xfov)
    depth = orc.apply_depth_scale(orc.decode_rgb_depth_frame(g[tag + "_depth_rgb"], 100, True), scale)
    unused, normals = io.edge_vertices(depth, K, True)
    assert np.array_equal(unused, g[tag + "_unused"]) and np.array_equal(normals, g[tag + "_removed_normals"])
    pts, ends = io.edge_points(depth, K, unused, normals)
    theta = None if conv == 0 else orc.convergence_angle(conv * scale, 0.063)
    M = orc.eye_pose("left", 0.063, theta) @ g[tag + "_transform"]
    pre, img, green, area = io.mask_before_inpaint(g[tag + "_left_image_u8"], g[tag + "_colour"], pts, ends, unused, M, K)
    assert np.array_equal(pre, g[tag + "_mask_pre_inpaint"]) and np.array_equal(area * 255, g[tag + "_infill_area"])
    assert np.array_equal(img, g[tag + "_image_final"])
    assert np.array_equal(io.finish_mask(pre, green, area), g[tag + "_mask_final"])
    assert len(unused) > 100 and (pre != 0).any()


@pytest.mark.parametrize("tag", ["plain", "posed"])
def test_normal_march_oracle_equals_reference(golden_dir, tag):
    """--do_basic_infill: the scalar float32 restatement against the reference's infill_using_normals output."""
    g = np.load(os.path.join(golden_dir, "infill_mask.npz"))
    hole = g[tag + "_hole_mask"]
    image = g[tag + "_left_image_u8"].copy()
    image[hole] = 0
    assert np.array_equal(io.normal_march_infill(image, hole, g[tag + "_mask_final"]), g[tag + "_image_basic_infill"])
