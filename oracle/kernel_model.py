"""float32, operation-for-operation NumPy model of the CUDA kernels' arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/mdvt_oracle.py).  The product computes geometry in float32
(one IEEE rounding per written operation, no FMA contraction: csrc is compiled with --fmad=false
and spells every operation with __fmul_rn/__fadd_rn/__fdiv_rn).  NumPy float32 +,-,*,/ round the
same way, so this model predicts the kernels' (u', v', z') *bit for bit*; feeding its output to
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
    """((m0*X + m1*Y) + m2*Z) + m3 in float32."""
    return ((m[0] * X + m[1] * Y) + m[2] * Z) + m[3]


def view_uvz_f32(depth_rgb, src, M, K_out, bit16=True):
    """project_splat_kernel's per-pixel arithmetic for one view; M is 3x4/4x4 float64 (rounded to float32
    like ViewSpec.to_c), K_out = (fx, fy, cx, cy)."""
    X, Y, Z = unproject_f32(depth_rgb, src, bit16)
    m = np.asarray(M, dtype=np.float64)[:3, :4].astype(f32)
    fx, fy, cx, cy = (f32(v) for v in K_out)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        Xv, Yv, Zv = affine_row(m[0], X, Y, Z), affine_row(m[1], X, Y, Z), affine_row(m[2], X, Y, Z)
        u = (fx * Xv) / Zv + cx
        v = (fy * Yv) / Zv + cy
    assert u.dtype == f32 and v.dtype == f32 and Zv.dtype == f32
    return u, v, Zv


def splat_ids_f32(u, v, z, out_w, out_h, near=orc.NEAR_PLANE):
    """The kernel's cull / round / bounds / atomicMin rule on float32 inputs == the oracle's rule
    (rintf is round-half-even; the float32 comparisons promote exactly)."""
    return orc.splat_ids(u.astype(np.float64), v.astype(np.float64), z.astype(np.float64), out_w, out_h, float(f32(near)))


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
