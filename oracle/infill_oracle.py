"""CPU oracle for the normals-coded infill mask: the edge test of the reference's mesh builder,
the edge points it paints into the disocclusion holes, and the mask image handed to the infill engines.

TEST INFRASTRUCTURE ONLY (same rules as oracle/mdvt_oracle.py).  A from-scratch NumPy restatement in grid form (no
triangle index arrays); pinned by tests/golden/infill_mask.npz, which oracle/make_infill_golden.py produced by running
the reference's own code (depth_map_tools.create_mesh_from_point_cloud through a fake Open3D shim, and the
stereo_rerender.py lines by line range).
"""
from __future__ import annotations

import math

import numpy as np

from . import mdvt_oracle as orc

EDGE_ANGLE_DEG = 89.0  # depth_map_tools.py:1192


def _cross(a, b):
    return np.stack((a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]), axis=-1)


def _dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def cell_triangles(vertices_hw3: np.ndarray):
    """Per grid cell (i, j), i < H-1, j < W-1: the two triangles of depth_map_tools.py:1243-1254,
    A = ((i,j), (i+1,j), (i+1,j+1)) and B = ((i,j), (i+1,j+1), (i,j+1)).  Returns (normal_A, normal_B, bad_A, bad_B):
    raw normals cross(v2 - v1, v3 - v1) and the 89-degree test against the view vector -centroid (:1283-1294)."""
    p00, p10 = vertices_hw3[:-1, :-1], vertices_hw3[1:, :-1]
    p11, p01 = vertices_hw3[1:, 1:], vertices_hw3[:-1, 1:]
    cos_limit = np.cos(np.radians(EDGE_ANGLE_DEG))
    out = []
    for v1, v2, v3 in ((p00, p10, p11), (p00, p11, p01)):
        n = _cross(v2 - v1, v3 - v1)
        view = -(v1 + v2 + v3) / 3.0
        cosine = _dot(n, view) / (np.sqrt(_dot(n, n)) * np.sqrt(_dot(view, view)) + 1e-15)
        out.append((n, cosine < cos_limit))
    (na, bad_a), (nb, bad_b) = out
    return na, nb, bad_a, bad_b


def edge_vertices(depth: np.ndarray, K: np.ndarray, of_by_one: bool = True):
    """get_mesh_from_depth_map(..., remove_edges=True, return_normals_of_removed=True) minus the mesh:
    (unused vertex indices ascending, their normals).  A vertex is unused when any triangle it belongs to fails
    the angle test (:1329-1335); its normal is the unit normal of the LAST triangle that lists it, in the order
    "all A triangles row-major, then all B triangles" (:1337-1361; unit normal := (1,1,1) where the area is 0)."""
    h, w = depth.shape
    verts = orc.unproject(depth, K, of_by_one).reshape(h, w, 3)
    na, nb, bad_a, bad_b = cell_triangles(verts)
    unused = np.zeros((h, w), dtype=bool)
    for bad, corners in ((bad_a, ((0, 0), (1, 0), (1, 1))), (bad_b, ((0, 0), (1, 1), (0, 1)))):
        for di, dj in corners:
            unused[di:h - 1 + di, dj:w - 1 + dj] |= bad

    def unit(n):
        length = np.sqrt((n * n).sum(axis=-1))[..., None]  # np.linalg.norm
        return np.divide(n, length, out=np.ones_like(n), where=length > 0)

    ua, ub = unit(na), unit(nb)
    vn = np.zeros((h, w, 3))
    vn[h - 1, 0] = ua[h - 2, 0]          # only A of cell (H-2, 0) lists the bottom-left vertex (as its 2nd corner)
    vn[h - 1, 1:] = ub[h - 2, :]         # bottom row: B of cell (H-2, j-1), 2nd corner
    vn[:h - 1, w - 1] = ub[:, w - 2]     # right column: B of cell (i, W-2), 3rd corner
    vn[:h - 1, :w - 1] = ub              # everywhere else: B of the vertex's own cell, 1st corner
    idx = np.flatnonzero(unused.reshape(-1))
    return idx, vn.reshape(-1, 3)[idx]


def edge_points(depth: np.ndarray, K: np.ndarray, unused: np.ndarray, normals: np.ndarray):
    """stereo_rerender.py:589-606: (edge_points, normal end points) in frame space.  The end points are built on the
    stretched (of_by_one) vertices BEFORE those are squeezed back by (W-1)/W, (H-1)/H -- kept as the reference has it."""
    h, w = depth.shape
    stretched = orc.unproject(depth, K, True)[unused]
    ends = normals + stretched
    pts = stretched.copy()
    pts[:, 0] *= (w - 1) / w
    pts[:, 1] *= (h - 1) / h
    return pts, ends


def mask_before_inpaint(image_u8: np.ndarray, colour: np.ndarray, pts, ends, unused, M, K_render, bg_rgb=(0, 255, 0)):
    """stereo_rerender.py:733-735,740-805,813-814 for one eye.  image_u8: the rendered eye with `bg_rgb` at holes.
    M: 4x4 frame -> eye camera.  Returns (mask image u8 before inpainting, eye image u8 with holes blacked and edge
    colours painted, bool hole mask, bool inpaint area)."""
    h, w = image_u8.shape[:2]
    hole = np.all(image_u8 == np.asarray(bg_rgb, dtype=np.uint8), axis=-1)
    mask = np.zeros((h, w, 3), dtype=np.float64)
    bg = np.asarray(bg_rgb, dtype=np.float64) / 255.0
    mask[hole] = bg
    image = image_u8.copy()
    image[hole] = 0
    for sel, val in (((slice(None), 0), (1.0, 0.5, 0.5)), ((slice(None), -1), (0.0, 0.5, 0.5)),
                     ((0, slice(None)), (0.5, 0.5, 0.0)), ((-1, slice(None)), (0.5, 0.5, 1.0))):   # :796-799
        line = mask[sel]
        line[np.all(line == bg, axis=-1)] = val
    if len(pts) > 1:
        q = orc.apply_pose(pts, M)
        q_end = orc.apply_pose(ends, M)
        k32 = np.asarray(K_render).astype(np.float32).astype(np.float64)
        u, v, z = orc.project(q, k32)
        ui, vi = np.round(u), np.round(v)
        with np.errstate(invalid="ignore"):
            ok = (ui >= 0) & (ui < w) & (vi >= 0) & (vi < h)
        order = np.argsort(z[ok], kind="stable")[::-1]                    # far -> near, the nearest is written last
        tx, ty = ui[ok].astype(int)[order], vi[ok].astype(int)[order]
        on_hole = hole[ty, tx]
        normal = (q_end - q)[ok][order][on_hole]
        normal = normal / np.linalg.norm(normal, axis=1, keepdims=True)
        mask[ty[on_hole], tx[on_hole]] = (normal + 1) / 2
        image[ty[on_hole], tx[on_hole]] = colour.reshape(-1, 3)[unused][ok][order][on_hole]
    green = np.all(mask == bg, axis=-1)
    area = green | np.all(mask == 0.0, axis=-1)
    return (mask * 255).astype(np.uint8), image, green, area


def finish_mask(mask_u8: np.ndarray, green: np.ndarray, area: np.ndarray) -> np.ndarray:
    """stereo_rerender.py:805-808,817: TELEA inpaint of everything that is not a coded normal, copied into the hole
    pixels only, then the black-ignoring 6x6 Gaussian (masked_blur, :114-153)."""
    import cv2

    filled = cv2.inpaint(mask_u8, (area * 255).astype(np.uint8), inpaintRadius=3, flags=cv2.INPAINT_TELEA)
    mask = mask_u8.astype(np.float64) / 255.0
    mask[green] = filled[green].astype(np.float32) / 255.0
    return (masked_blur((mask * 255).astype(np.uint8)).astype(np.float32) / 255.0 * 255).astype(np.uint8)


def masked_blur(img: np.ndarray, ksize=(6, 6), sigma=0) -> np.ndarray:
    """stereo_rerender.py:114-153: Gaussian blur whose kernel weights skip pure-black pixels."""
    import cv2

    g = cv2.getGaussianKernel(ksize[0], sigma)
    kernel = g @ g.T
    black = np.all(img == 0, axis=2)
    valid = (~black).astype(np.float32)
    acc = cv2.filter2D(img.astype(np.float32), -1, kernel, borderType=cv2.BORDER_ISOLATED)
    weight = cv2.filter2D(valid, -1, kernel, borderType=cv2.BORDER_ISOLATED)
    out = acc / np.where(weight[..., None] == 0, 1.0, weight[..., None])
    out[weight == 0] = 0
    out[black] = 0
    return np.clip(out, 0, 255).astype(np.uint8)


def normal_march_infill(image_u8: np.ndarray, hole: np.ndarray, mask_final_u8: np.ndarray, max_steps: int = 400) -> np.ndarray:
    """`--do_basic_infill` (stereo_rerender.py:155-240,810-812) per pixel, scalar float32: every hole pixel marches
    along the XY direction coded in the final mask image ((m/255)*2-1, normalised) until it leaves the hole, then takes
    the colour two steps / one step / zero steps further on, whichever is first inside the frame and not a hole.
    Pure-Python loop: small frames only."""
    f32 = np.float32
    h, w = hole.shape
    out = image_u8.copy()
    for y, x in zip(*np.nonzero(hole)):
        m = mask_final_u8[y, x].astype(f32) / f32(255.0)
        dx, dy = m[0] * f32(2) - f32(1), m[1] * f32(2) - f32(1)
        norm = np.sqrt(dx * dx + dy * dy)
        if not norm > f32(1e-6):
            continue
        dx, dy = dx / norm, dy / norm
        for t in range(1, max_steps + 1):
            xi, yi = int(np.rint(f32(x) + dx * f32(t))), int(np.rint(f32(y) + dy * f32(t)))
            if not (0 <= xi < w and 0 <= yi < h):
                break
            if hole[yi, xi]:
                continue
            for dt in (2, 1, 0):
                x2, y2 = int(np.rint(f32(x) + dx * f32(t + dt))), int(np.rint(f32(y) + dy * f32(t + dt)))
                if 0 <= x2 < w and 0 <= y2 < h and not hole[y2, x2]:
                    out[y, x] = image_u8[y2, x2]
                    break
            break
    return out
