"""float32, operation-for-operation NumPy model of the CUDA kernels' arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/mdvt_oracle.py).  The product computes geometry in float32
with every operation spelled as an explicit intrinsic (__fmul_rn / __fadd_rn / __fdiv_rn / __fmaf_rn;
csrc is compiled with --fmad=false, so nothing is contracted behind the source's back).  NumPy
float32 +,-,*,/ round the same way and `fma32` below is an exact float32 fused multiply-add,
so this model predicts the kernels' (u', v', z') *bit for bit*; feeding its output to
the oracle's integer stage (`mdvt_oracle.splat_ids` / `resolve`) then predicts every output
byte.  Tests use it in two directions:

  * CUDA == model, bit-exact  (index / byte parity of the integer work)
  * model ~= float64 oracle within 1e-4 relative (the float tolerance BASELINE.json states),
    and winners identical except where the float64 (u', v') sits within 1e-3 px of a .5
    rounding boundary or two candidates' z' differ by < 1e-5 relative (counted, reported).
"""
from __future__ import annotations

import numpy as np

from . import mdvt_oracle as orc

f32 = np.float32


def source_constants(width, height, K, max_depth, decoder="D1", depth_scale=1.0, of_by_one=False):
    """Mirror of metric_depth_video_toolbox_b200.ops.make_source (float32 roundings of the float64 inputs)."""
    dec = f32(float(max_depth) / orc.FULL_SCALE) if decoder == "D1" else f32(orc.FULL_SCALE / max_depth)
    return dict(width=width, height=height, decoder=decoder, dec_const=dec, depth_scale=f32(depth_scale),
                fx=f32(K[0][0]), fy=f32(K[1][1]), cx=f32(K[0][2]), cy=f32(K[1][2]),
                sx=f32((width + 1) / width) if of_by_one else f32(1.0),
                sy=f32((height + 1) / height) if of_by_one else f32(1.0))


def depth_f32(depth_rgb, src, bit16=True):
    """depth_of<> then the depth_scale multiply (csrc/mdvt_common.cuh)."""
    codes = orc.decode_codes(depth_rgb, bit16, src["decoder"]).astype(f32)
    d = codes * src["dec_const"] if src["decoder"] == "D1" else codes / src["dec_const"]
    return (d * src["depth_scale"]).astype(f32)


def unproject_f32(depth_rgb, src, bit16=True):
    """unproject_px: X = ((col*sx - cx) * z) / fx, float32 at every step."""
    h, w = src["height"], src["width"]
    z = depth_f32(depth_rgb, src, bit16)
    col, row = np.meshgrid(np.arange(w, dtype=f32), np.arange(h, dtype=f32))
    xg, yg = col * src["sx"], row * src["sy"]
    X = ((xg - src["cx"]) * z) / src["fx"]
    Y = ((yg - src["cy"]) * z) / src["fy"]
    return X.reshape(-1), Y.reshape(-1), z.reshape(-1)


def affine_row(m, X, Y, Z):
    """((m0*X + m1*Y) + m2*Z) + m3 in float32 (unproject_f32_kernel's optional pose; splat_points_kernel)."""
    return ((m[0] * X + m[1] * Y) + m[2] * Z) + m[3]


def fma32(a, b, c):
    """fmaf(a, b, c): the exactly rounded float32 of a*b + c, element-wise (what __fmaf_rn / FFMA computes).

    a*b is exact in float64 (24 + 24 significant bits).  The float64 sum p + c rounds once more, which could
    double-round at a float32 tie, so the sum is rounded TO ODD instead (TwoSum gives the exact rounding error;
    an inexact sum with an even last bit moves to its odd neighbour on the error's side): rounding a
    round-to-odd float64 (53 bits >= 24 + 2) to float32 equals rounding the exact value."""
    a, b, c = (np.asarray(x, dtype=np.float32) for x in (a, b, c))
    p = a.astype(np.float64) * b.astype(np.float64)
    c64 = c.astype(np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        s = p + c64
        bb = s - p
        err = (p - (s - bb)) + (c64 - bb)
    bits = np.atleast_1d(s).view(np.int64).copy()
    fix = np.atleast_1d(np.isfinite(s) & (err != 0) & np.isfinite(err)) & ((bits & 1) == 0)
    away = np.atleast_1d((err > 0) == (s > 0))   # the exact value lies further from zero than s
    bits[fix & away] += 1                          # sign-magnitude: +1 grows the magnitude
    bits[fix & ~away] -= 1
    out = bits.view(np.float64).reshape(np.shape(s))
    with np.errstate(over="ignore", invalid="ignore"):
        return out.astype(np.float32)


def ray_view(src, M, K_out):
    """make_ray_view (csrc/mdvt_common.cuh): the twelve float32 coefficients of one view, evaluated in float64 with
    the same operations in the same order.  M: 3x4 / 4x4 (rounded to float32 first, like ViewSpec.to_c);
    K_out = (fx', fy', cx', cy')."""
    f64 = np.float64
    m = np.asarray(M, dtype=f64)[:3, :4].astype(f32).astype(f64)
    kfx, kfy, kcx, kcy = (f64(f32(v)) for v in K_out)
    fx, fy, cx, cy, sx, sy = (f64(src[k]) for k in ("fx", "fy", "cx", "cy", "sx", "sy"))
    ax, bx, ay, by = sx / fx, -cx / fx, sy / fy, -cy / fy
    P = np.stack([kfx * m[0] + kcx * m[2], kfy * m[1] + kcy * m[2], m[2]])
    A = (P[:, 0] * ax).astype(f32)
    B = (P[:, 1] * ay).astype(f32)
    C = ((P[:, 0] * bx + P[:, 1] * by) + P[:, 2]).astype(f32)
    T = P[:, 3].astype(f32)
    return A, B, C, T


def view_uvz_f32(depth_rgb, src, M, K_out, bit16=True):
    """project_splat_kernel's / stereo_conv_rows_kernel's per-pixel arithmetic for one view (ray form):
    r_c = fma(B_c, row, fma(A_c, col, C_c)); N_c = fma(z, r_c, T_c); u = N_u / Zv, v = N_v / Zv, Zv = N_z."""
    h, w = src["height"], src["width"]
    z = depth_f32(depth_rgb, src, bit16).reshape(h, w)
    A, B, C, T = ray_view(src, M, K_out)
    col = np.arange(w, dtype=f32)[None, :]
    row = np.arange(h, dtype=f32)[:, None]
    n = []
    for c in range(3):
        cj = fma32(A[c], col, C[c])
        r = fma32(B[c], row, np.broadcast_to(cj, (h, w)))
        n.append(fma32(z, r, T[c]))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        u = n[0] / n[2]
        v = n[1] / n[2]
    assert u.dtype == f32 and v.dtype == f32 and n[2].dtype == f32
    return u.reshape(-1), v.reshape(-1), n[2].reshape(-1)


def splat_ids_f32(u, v, z, out_w, out_h, near=orc.NEAR_PLANE, tie=None):
    """The kernel's cull / round / bounds / atomicMin rule on float32 inputs == the oracle's rule
    (rintf is round-half-even; the float32 comparisons promote exactly).  tie: the packed colours for the
    colour-keyed frame loops (key = Zv bits << 32 | 0x00BBGGRR), None for the index-keyed primitives."""
    return orc.splat_ids(u.astype(np.float64), v.astype(np.float64), z.astype(np.float64), out_w, out_h, float(f32(near)), tie)


def render_views_f32(depth_rgb, colour, src, views, out_w, out_h, bg_rgb=(0, 0, 0), fill_rgb=(0, 0, 0), bg_collide=False, near=orc.NEAR_PLANE):
    """mdvt_render_views / mdvt_stereo_conv_rows / mdvt_novel_view_frames for one frame: every output byte and the float32
    depth planes.  views: [(M, (fx, fy, cx, cy)), ...].  Returns (image (out_h, n*out_w, 3) u8, mask (out_h, n*out_w) u8,
    depth (out_h, n*out_w) f32, [ids per view])."""
    imgs, masks, depths, idl = [], [], [], []
    tie = orc.pack_colour(colour)
    for M, K_out in views:
        u, v, z = view_uvz_f32(depth_rgb, src, M, K_out)
        ids = splat_ids_f32(u, v, z, out_w, out_h, near, tie)
        img, mask = orc.resolve(ids, colour, bg_rgb, fill_rgb, bg_collide)
        imgs.append(img)
        masks.append(mask)
        depths.append(orc.zbuffer_depth(ids, z))
        idl.append(ids)
    return np.concatenate(imgs, axis=1), np.concatenate(masks, axis=1), np.concatenate(depths, axis=1), idl


def stereo_rows_f32(depth_rgb, colour, consts, bg_rgb=(0, 0, 0), fill_rgb=(0, 0, 0), bg_collide=False):
    """stereo_rows_kernel for one frame.  consts = (dec_const, depth_scale, fx_half_ipd, near) float32.
    Returns (sbs (H,2W,3) u8, mask (H,2W) u8, (ids_left, ids_right))."""
    dec, scale, fxs, near = (f32(c) for c in consts)
    h, w = depth_rgb.shape[:2]
    c16 = (depth_rgb[..., 0].astype(np.uint32) << 8) | depth_rgb[..., 2].astype(np.uint32)
    z = ((c16 << 16).astype(f32) * dec) * scale
    with np.errstate(divide="ignore"):
        d = fxs / z
    # target column = round-half-even of the EXACT sum j +- d, rounded once (the kernels add d to the exactly
    # representable 1.5*2^23 + j, whose float32 ulp is 1); float64 holds j + d exactly whenever it can tie
    col = np.broadcast_to(np.arange(w, dtype=np.float64), (h, w))
    row = np.broadcast_to(np.arange(h, dtype=np.float64)[:, None], (h, w)).reshape(-1)
    outs = []
    for sign in (+1, -1):
        u = col + sign * d.astype(np.float64)
        ids = orc.splat_ids(u.reshape(-1), row, z.reshape(-1).astype(np.float64), w, h, float(near))
        img, mask = orc.resolve(ids, colour, bg_rgb, fill_rgb, bg_collide)
        outs.append((img, mask, ids))
    sbs = np.concatenate([outs[0][0], outs[1][0]], axis=1)
    mask = np.concatenate([outs[0][1], outs[1][1]], axis=1)
    return sbs, mask, (outs[0][2], outs[1][2])
