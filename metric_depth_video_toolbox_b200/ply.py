"""Legacy-Open3D-style binary PLY point clouds (what `o3d.io.write_point_cloud` produces for the
convert script's --save_ply, convert_metric_depth_video_to_other_format.py:743-749): little-endian,
`double x y z` + `uchar red green blue`, 27 bytes per point.  Host file I/O only."""
from __future__ import annotations

import numpy as np

_VERTEX = np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])


def write_point_cloud(path: str, xyz: np.ndarray, rgb_u8: np.ndarray) -> None:
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    rgb = np.asarray(rgb_u8, dtype=np.uint8).reshape(-1, 3)
    if len(xyz) != len(rgb):
        raise ValueError("one colour per point is required")
    rec = np.empty(len(xyz), dtype=_VERTEX)
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rec["red"], rec["green"], rec["blue"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    header = ("ply\nformat binary_little_endian 1.0\ncomment Created by Open3D\n"
              f"element vertex {len(xyz)}\nproperty double x\nproperty double y\nproperty double z\n"
              "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        rec.tofile(fh)


def read_point_cloud(path: str):
    with open(path, "rb") as fh:
        n = None
        while True:
            line = fh.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
        rec = np.fromfile(fh, dtype=_VERTEX, count=n)
    return np.stack((rec["x"], rec["y"], rec["z"]), axis=-1), np.stack((rec["red"], rec["green"], rec["blue"]), axis=-1)
