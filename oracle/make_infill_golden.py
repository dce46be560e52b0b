"""Golden vectors for the normals-coded infill mask, made by RUNNING the reference:
`depth_map_tools.get_mesh_from_depth_map(remove_edges=True, return_normals_of_removed=True)` through the fake Open3D
shim, then the edge-point / mask-assembly lines of stereo_rerender.py (exec'd by line range from the read-only
checkout) on a left-eye image rendered by the point-splat oracle.

    python oracle/make_infill_golden.py      # needs /root/reference; writes tests/golden/infill_mask.npz

TEST INFRASTRUCTURE ONLY.  Stores inputs and reference OUTPUTS, no reference source.
"""
from __future__ import annotations

import argparse
import copy
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import mdvt_oracle as orc  # noqa: E402
from oracle import ref_bridge  # noqa: E402
from metric_depth_video_toolbox_b200.synth import SyntheticClip  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "infill_mask.npz")


def run_case(dfh, dmt, sr, tag, w, h, xfov, conv_depth, transform, out):
    depth_rgb, colour = SyntheticClip(w, h, 3, zero_fraction=0.003).frame(1)
    K = dmt.compute_camera_matrix(xfov, None, w, h)
    scale = orc.master_fov_depth_scale(45.0, xfov)
    depth = dfh.decode_rgb_depth_frame(depth_rgb, 100, True)
    depth *= scale                                                       # stereo_rerender.py:541
    mesh, unused, normals = dmt.get_mesh_from_depth_map(depth, K, colour, None, remove_edges=True, of_by_one=True,
                                                        return_normals_of_removed=True)   # :583
    out[f"{tag}_depth_rgb"], out[f"{tag}_colour"] = depth_rgb, colour
    out[f"{tag}_params"] = np.array([w, h, xfov, 0.0 if conv_depth is None else conv_depth])
    out[f"{tag}_transform"] = np.eye(4) if transform is None else transform
    out[f"{tag}_unused"], out[f"{tag}_removed_normals"] = np.asarray(unused), np.asarray(normals)
    ipd = 0.063
    args = argparse.Namespace(dont_place_points_in_edges=False, green_and_black_infill_mask=False, do_basic_infill=False, pupillary_distance=63)
    ns = {"np": np, "cv2": cv2, "args": args, "remove_edges": True, "mesh": mesh, "unused_indices": unused, "removed_normals": normals,
          "frame_width": w, "frame_height": h, "depth_map_tools": dmt, "edge_pcd": None, "masked_blur": sr.masked_blur,
          "infill_using_normals": sr.infill_using_normals, "render_cam_matrix": K, "left_shift": -ipd / 2,
          "bg_color": np.array([0.0, 1.0, 0.0]), "infill_mask_video": object(), "draw_mesh": copy.deepcopy(mesh)}
    ref_bridge.exec_lines("stereo_rerender.py", 589, 606, ns)            # edge points, world-space normal end points
    M = np.eye(4)
    if transform is not None:
        ns["transformations"], ns["transform_to_zero"] = [transform], transform
        ref_bridge.exec_lines("stereo_rerender.py", 615, 619, ns)
        M = transform
    theta = None
    ns["convergence_distance"] = None
    if conv_depth is not None:
        conv = conv_depth * scale                                        # :716
        theta = sr.convergence_angle(conv, ipd)                          # :719
        ns["convergence_distance"] = conv
        ns["convergence_rotation_minus"] = mesh.get_rotation_matrix_from_xyz((0, -theta, 0))
    ref_bridge.exec_lines("stereo_rerender.py", 723, 735, ns)            # eye pose on mesh + edge clouds, projection
    M = orc.eye_pose("left", ipd, theta) @ M
    img, _, _ = orc.render_view(depth_rgb, colour, 100, K, M, depth_scale=scale, bg_rgb=(0, 255, 0), hole_fill=(0, 255, 0))
    ns["left_image"] = (img.astype(np.float32) / np.float32(255.0))      # what render() hands back (:738)
    out[f"{tag}_left_image_u8"] = img
    ref_bridge.exec_lines("stereo_rerender.py", 740, 805, ns)            # hole mask, border + edge normals, inpaint area
    out[f"{tag}_mask_pre_inpaint"] = (ns["left_img_mask"] * 255).astype("uint8")
    out[f"{tag}_infill_area"] = ns["infill_area_mask"]
    ref_bridge.exec_lines("stereo_rerender.py", 806, 808, ns)            # TELEA + masked blur
    basic = dict(ns)                                                     # --do_basic_infill variant of :810-819 (normal-march infill)
    basic["args"] = argparse.Namespace(**{**vars(args), "do_basic_infill": True})
    basic["left_image"], basic["left_img_mask"] = ns["left_image"].copy(), ns["left_img_mask"].copy()
    ref_bridge.exec_lines("stereo_rerender.py", 810, 819, basic)
    out[f"{tag}_image_basic_infill"] = basic["left_image"]
    out[f"{tag}_hole_mask"] = ns["bg_mask"]
    ref_bridge.exec_lines("stereo_rerender.py", 810, 819, ns)            # edge colours into the image, u8 conversions
    out[f"{tag}_mask_final"], out[f"{tag}_image_final"] = ns["left_img_mask"], ns["left_image"]
    print(tag, "edge vertices", len(unused), "painted", int(ns["mask"].sum()), "holes", int(ns["bg_mask"].sum()))


def main():
    if not ref_bridge.available():
        raise SystemExit(f"reference checkout not found at {ref_bridge.REFERENCE_ROOT}")
    dfh, dmt, sr = (ref_bridge.load(m) for m in ("depth_frames_helper", "depth_map_tools", "stereo_rerender"))
    out = {}
    run_case(dfh, dmt, sr, "plain", 96, 64, 60.0, None, None, out)
    T = np.eye(4)
    T[:3, :3] = orc.rot_y(0.01)
    T[:3, 3] = (0.02, -0.01, 0.05)
    run_case(dfh, dmt, sr, "posed", 96, 64, 70.0, 4.0, T, out)
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
