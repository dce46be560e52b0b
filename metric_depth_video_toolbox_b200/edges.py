"""`remove_edges=True` for the drop-in get_mesh_from_depth_map: which vertices the reference's mesh builder would take
out (depth_map_tools.py:1243-1376) and the normals it returns for them, from the GPU edge kernels."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def edge_vertices(mesh):
    """(unused vertex indices ascending (int64), normals_of_removed (k, 3) float64) for a DepthMesh."""
    flags, normals = ops.edge_vertices(mesh.depth, mesh.source(), mesh.K)
    idx = torch.nonzero(flags.reshape(-1), as_tuple=False).reshape(-1)
    return idx.cpu().numpy().astype(np.int64), normals.reshape(-1, 3)[idx].cpu().numpy()
