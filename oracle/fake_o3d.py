"""A minimal stand-in for the few Open3D classes the reference's mesh-building code touches, so that
`depth_map_tools.create_mesh_from_point_cloud` / `pts_2_pcd` and the edge-point lines of stereo_rerender.py can be
RUN (not copied) when generating golden vectors.  TEST INFRASTRUCTURE ONLY: plain NumPy containers with Open3D's
in-place semantics (np.asarray(mesh.vertices) aliases the stored array; transform / rotate / translate mutate it)."""
from __future__ import annotations

import types

import numpy as np


class _Geometry:
    def _pts(self):
        return self.vertices if hasattr(self, "vertices") else self.points

    def transform(self, T):
        T = np.asarray(T, dtype=np.float64)
        p = self._pts()
        if len(p):
            p[:] = p @ T[:3, :3].T + T[:3, 3]
        return self

    def rotate(self, R, center=(0, 0, 0)):
        c = np.asarray(center, dtype=np.float64)
        p = self._pts()
        p[:] = (p - c) @ np.asarray(R, dtype=np.float64).T + c
        return self

    def translate(self, t, relative=True):
        self._pts()[:] += np.asarray(t, dtype=np.float64)
        return self

    def get_center(self):
        return self._pts().mean(axis=0)

    @staticmethod
    def get_rotation_matrix_from_xyz(angles):
        a, b, c = angles
        rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
        ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
        return rx @ ry @ rz


class TriangleMesh(_Geometry):
    def __init__(self):
        self.vertices = np.zeros((0, 3))
        self.triangles = np.zeros((0, 3), dtype=np.int64)
        self.vertex_colors = np.zeros((0, 3))


class PointCloud(_Geometry):
    def __init__(self):
        self.points = np.zeros((0, 3))
        self.colors = np.zeros((0, 3))
        self.normals = np.zeros((0, 3))


def module() -> types.ModuleType:
    o3d = types.ModuleType("open3d")
    o3d.geometry = types.SimpleNamespace(TriangleMesh=TriangleMesh, PointCloud=PointCloud)
    o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: np.array(a, dtype=np.float64),
                                        Vector3iVector=lambda a: np.array(a, dtype=np.int64))
    return o3d
