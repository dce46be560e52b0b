// Generic novel-view path: K1+K2 (decode -> unproject -> pose -> project -> 64-bit atomicMin splat) and
// K3 (resolve: gather colour by winning source index, hole mask, depth plane, z-buffer reset).
//
// The z-buffer is a u64 plane per view: (float_bits(z') << 32) | source_index.  z' > near > 0, so the
// float bit pattern orders like the value and one unsigned atomicMin implements "nearest wins, ties ->
// lowest source index" deterministically.  1080p stereo = 33 MB: resident in the 126 MB L2 (ncu: 0.5 MB of DRAM reads
// per splat).  One 4K view = 66 MB does NOT stay resident next to the streams (ncu: 43 % L2 hit rate; the two-die L2
// holds about half of its nominal size for one working set), which is what the touched-segment flags below are for.
#include <cstdlib>
#include <mutex>

#include "mdvt_common.cuh"

namespace mdvt {

constexpr int kThreads = 256;
constexpr int kMaxViews = 4;
constexpr int kSplatThreads = 128;  // divides 640 / 1280 / 1920 / 3840: no idle tail block per row
constexpr float kRoundMagic = 12582912.0f;     // 1.5 * 2^23
constexpr int kRoundMagicBits = 0x4B400000;
// "Touched" flags: one byte per 64 consecutive slots of the z-buffer plane, set by the splat next to its RED (single view only).
// The resolve skips the z-buffer load AND the re-arm store of untouched segments: with the 62 % holes of the config-3
// camera, z-buffer traffic drops from 2 x 66 MB to 2 x 25 MB per 4K frame and the touched part stays L2-resident.
constexpr int kSegShift = 6;

static int g_grid_pct() {  // development switch: MDVT_GRID_PCT = share of the resident CTA slots a frame-loop kernel asks for (default 100)
    static const int pct = getenv("MDVT_GRID_PCT") ? atoi(getenv("MDVT_GRID_PCT")) : 100;
    return pct < 10 ? 10 : (pct > 100 ? 100 : pct);
}

struct ViewPack {
    mdvt_view v[kMaxViews];
    int n;
};
struct RayPack {  // the views of one frame in ray form (mdvt_common.cuh)
    RayView v[kMaxViews];
    int n;
};

// One source pixel through every view in ray form.  cj[k][c] = fma(A_c, col, C_c) is hoisted per column by the caller,
// fi = float(row).  Divisions are correctly rounded (== the float32 model's `/`): the reciprocal of Zv is refined once
// per view and shared by u and v.  All index arithmetic is 32-bit (the entry point checks that source and target
// planes have < 2^31 pixels).
// EPOCH (the colour-keyed frame loops): key = epoch << 56 | float_bits(Zv) << 25 | colour.  Frames count the epoch DOWN, so
// a key written for the current frame is smaller than anything an earlier frame left in the plane: stale slots lose every
// atomicMin and the resolve recognises them by their epoch byte -- the z-buffer is never re-armed between frames (that
// pass wrote 8 bytes per target pixel and view; the planes are cleared once per call instead).
template <bool UVZ, bool TOUCHED, bool EPOCH>
__device__ __forceinline__ void splat_pixel(uint32_t p, const float (&cj)[kMaxViews][3], float fi, float z, const RayPack &rays,
                                            float near_plane, int out_w, uint32_t out_n, uint32_t out_h, uint32_t payload, uint32_t n,
                                            unsigned long long *__restrict__ zbuf, float *__restrict__ out_uvz, uint64_t keep,
                                            uint8_t *__restrict__ touched, uint32_t epoch_hi, uint32_t lanes) {
#pragma unroll
    for (int k = 0; k < kMaxViews; ++k) {
        if (k < rays.n) {
            const RayView &rv = rays.v[k];
            const float nu = __fmaf_rn(z, __fmaf_rn(rv.B[0], fi, cj[k][0]), rv.T[0]);
            const float nv = __fmaf_rn(z, __fmaf_rn(rv.B[1], fi, cj[k][1]), rv.T[1]);
            const float Zv = __fmaf_rn(z, __fmaf_rn(rv.B[2], fi, cj[k][2]), rv.T[2]);
            float u, v;
            if (UVZ) {  // parity output: IEEE division also where Zv is zero / negative / tiny
                u = __fdiv_rn(nu, Zv);
                v = __fdiv_rn(nv, Zv);
                float *o = out_uvz + ((int64_t)k * n + p) * 3;
                o[0] = u; o[1] = v; o[2] = Zv;
            } else {
                const float rz = rcp_refined(Zv);
                u = div_rn_by(nu, Zv, rz);
                v = div_rn_by(nv, Zv, rz);
            }
            // Round half to even, like np.round, and test the bounds without FRND / F2I and with two compares: adding
            // 1.5 * 2^23 rounds to an integer (the sum's ulp is 1) for |u| < 2^22 and leaves rint(u) in the mantissa, so
            // bits(sum) - bits(1.5 * 2^23) is rint(u) as a signed integer and ONE unsigned compare tests 0 <= rint(u) < W.
            // |u| >= 2^22, +-inf and NaN give differences >= 2^22 or "negative" ones, i.e. huge unsigned values: rejected
            // (out_w, out_h < 2^22 is checked at the entry points).  Zv > near is false for NaN.
            const uint32_t ui = (uint32_t)(__float_as_int(__fadd_rn(u, kRoundMagic)) - kRoundMagicBits);
            const uint32_t vi = (uint32_t)(__float_as_int(__fadd_rn(v, kRoundMagic)) - kRoundMagicBits);
            const bool ok = Zv > near_plane && ui < (uint32_t)out_w && vi < out_h;
            const uint32_t t = (uint32_t)k * out_n + vi * (uint32_t)out_w + ui;
            const uint32_t zb = __float_as_uint(Zv);  // positive, finite: < 2^31 where ok
            const unsigned long long key = EPOCH ? (((unsigned long long)(epoch_hi | (zb >> 7)) << 32) | ((zb << 25) | payload))
                                                 : (((unsigned long long)zb << 32) | payload);
            // (measured and dropped: letting a lane stay silent when a neighbouring lane holds a smaller key for the same slot -- the
            //  novel view folds 2.6 source pixels onto a drawn target pixel -- costs 17 instructions per pixel and buys nothing: the
            //  L2 processes a warp's REDs per sector, 40.4 vs 41.0 us per 4K frame, profiles/r02_novel_4k_v7_dedup_brief.txt)
            if (ok) {
                red_min_u64_keep(zbuf + t, key, keep);
                if (TOUCHED) touched[t >> kSegShift] = 1;  // flat 64-slot segments of the plane (single view: t < out_n)
            }
        }
    }
}

// 2-D launch: blockIdx.x tiles the columns, blockIdx.y strides over the rows, so (col, row) need no division and
// the lanes of a warp cover 32 consecutive source pixels (-> mostly consecutive z-buffer slots: few L2 sectors per
// 64-bit RED).  A 4-pixels-per-thread variant with word loads was measured SLOWER (29 vs 17 us at 1080p x 2 views):
// its lanes hit every fourth slot and each RED touches 4x the sectors -- the atomics, not the loads, matter here.
// Rows are taken in batches of kSplatRows with all source loads issued before the first pixel is pushed through the
// cameras (ncu v1: the kernel sat on the scoreboard of its own byte loads at 53 % issue-active), and the grid is sized
// from the occupancy API so that every CTA is resident (v1 launched 2370 CTAs onto 1924 slots: a second, mostly
// empty wave).  DEVVIEW: the single camera is read from device memory (written by the look-at kernel of the same
// stream, mdvt_novel_view_frames) instead of the kernel parameters, and brought into ray form by every thread.
constexpr int kSplatRows = 4;

// CKEY: the low key word is the source pixel's COLOUR (0x00BBGGRR) instead of its index -- the frame loops
// (mdvt_render_views, mdvt_novel_view_frames): the z-buffer then IS the image and the resolve is a streaming pass with
// no gather.  Nearest Zv still wins; among candidates with bit-identical Zv the smallest packed colour wins (the
// index-keyed primitives take the lowest source index; the reference leaves ties to an unstable argsort).
template <int DECODER, bool BIT16, bool DEVVIEW, bool CKEY>
__global__ void __launch_bounds__(kSplatThreads)
    project_splat_kernel(const void *__restrict__ rgb, const uint8_t *__restrict__ colour, int width, int height, float dec_const,
                         float depth_scale, SourceCam cam, RayPack rays_param, const mdvt_view *__restrict__ view_dev, float near_plane,
                         int out_w, int out_h, uint32_t id_offset, unsigned long long *__restrict__ zbuf, float *__restrict__ out_uvz,
                         uint8_t *__restrict__ touched, uint32_t epoch_hi) {
    const uint32_t out_n = (uint32_t)out_w * (uint32_t)out_h, n = (uint32_t)width * (uint32_t)height;
    const int col = blockIdx.x * kSplatThreads + threadIdx.x;
    if (col >= width) return;
    const uint32_t lanes = __activemask();  // the lanes of this warp that own a column (all 32 except in the last column block)
    RayPack local;
    if (DEVVIEW) {
        local.n = 1;
        mdvt_view vw;
        const float4 *v4 = reinterpret_cast<const float4 *>(view_dev);
        float4 *l4 = reinterpret_cast<float4 *>(&vw);
#pragma unroll
        for (int k = 0; k < 4; ++k) l4[k] = __ldg(v4 + k);
        local.v[0] = make_ray_view(cam.fx, cam.fy, cam.cx, cam.cy, cam.sx, cam.sy, vw.M, vw.fx, vw.fy, vw.cx, vw.cy);
    }
    const RayPack &rays = DEVVIEW ? local : rays_param;
    float cj[kMaxViews][3];
    const float fj = __int2float_rn(col);
#pragma unroll
    for (int k = 0; k < kMaxViews; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) cj[k][c] = k < rays.n ? __fmaf_rn(rays.v[k].A[c], fj, rays.v[k].C[c]) : 0.0f;
    const uint64_t keep = l2_keep_policy();
    const int stride = gridDim.y;
    for (int row0 = blockIdx.y; row0 < height; row0 += stride * kSplatRows) {
        float z[kSplatRows];
        uint32_t pay[kSplatRows];
#pragma unroll
        for (int k = 0; k < kSplatRows; ++k) {
            const int row = row0 + k * stride;
            z[k] = 0.0f;
            pay[k] = 0u;
            if (row < height) {
                const uint32_t p = (uint32_t)row * (uint32_t)width + (uint32_t)col;
                z[k] = source_depth<DECODER, BIT16>(rgb, p, dec_const);
                if (CKEY) {
                    const uint8_t *c = colour + (size_t)p * 3;
                    pay[k] = (uint32_t)__ldg(c) | ((uint32_t)__ldg(c + 1) << 8) | ((uint32_t)__ldg(c + 2) << 16);
                } else {
                    pay[k] = id_offset + p;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kSplatRows; ++k) {
            const int row = row0 + k * stride;
            if (row < height) {
                const uint32_t p = (uint32_t)row * (uint32_t)width + (uint32_t)col;
                const float zs = __fmul_rn(z[k], depth_scale), fi = __int2float_rn(row);
                // DEVVIEW: the novel-view frame loop -- touched flags always on, never a (u, v, z) dump
                if (DEVVIEW)
                    splat_pixel<false, true, true>(p, cj, fi, zs, rays, near_plane, out_w, out_n, (uint32_t)out_h, pay[k], n, zbuf, nullptr, keep,
                                                   touched, epoch_hi, lanes);
                else if (!CKEY && out_uvz)
                    splat_pixel<true, false, false>(p, cj, fi, zs, rays, near_plane, out_w, out_n, (uint32_t)out_h, pay[k], n, zbuf, out_uvz, keep,
                                                    nullptr, 0u, lanes);
                else
                    splat_pixel<false, false, CKEY>(p, cj, fi, zs, rays, near_plane, out_w, out_n, (uint32_t)out_h, pay[k], n, zbuf, nullptr, keep,
                                                    nullptr, epoch_hi, lanes);
            }
        }
    }
}

// Explicit points (N, 3) float32 in the frame space of the views' M: the reference's point painter
// (stereo_rerender.py:746-755,814) and render() of point clouds (background cloud, edge points).
__global__ void __launch_bounds__(kThreads)
    splat_points_kernel(const float *__restrict__ xyz, int64_t n, ViewPack views, float near_plane, int out_w, int out_h, uint32_t id_offset,
                        unsigned long long *__restrict__ zbuf) {
    const int64_t out_n = (int64_t)out_w * out_h;
    const float u_max = (float)(out_w - 1), v_max = (float)(out_h - 1);
    const uint64_t keep = l2_keep_policy();
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        const float X = __ldg(xyz + p * 3), Y = __ldg(xyz + p * 3 + 1), Z = __ldg(xyz + p * 3 + 2);
#pragma unroll
        for (int k = 0; k < kMaxViews; ++k) {
            if (k < views.n) {
                const mdvt_view &vw = views.v[k];
                const float Xv = affine_row(vw.M, X, Y, Z);
                const float Yv = affine_row(vw.M + 4, X, Y, Z);
                const float Zv = affine_row(vw.M + 8, X, Y, Z);
                const float ur = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(vw.fx, Xv), Zv), vw.cx));
                const float vr = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(vw.fy, Yv), Zv), vw.cy));
                if (Zv > near_plane && ur >= 0.0f && ur <= u_max && vr >= 0.0f && vr <= v_max) {
                    const int64_t t = (int64_t)(int)vr * out_w + (int)ur;
                    const unsigned long long key = ((unsigned long long)__float_as_uint(Zv) << 32) | (id_offset + (uint32_t)p);
                    red_min_u64_keep(zbuf + (int64_t)k * out_n + t, key, keep);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) zbuf_clear_kernel(unsigned long long *__restrict__ zbuf, int64_t n) {
    const uint64_t keep = l2_keep_policy();
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        st_u64_keep(zbuf + i, MDVT_ZBUF_EMPTY, keep);
}

__device__ __forceinline__ uint32_t gather_rgb(const uint8_t *__restrict__ colour, uint32_t id) {
    const uint8_t *c = colour + (int64_t)id * 3;
    return (uint32_t)__ldg(c) | ((uint32_t)__ldg(c + 1) << 8) | ((uint32_t)__ldg(c + 2) << 16);
}

// One thread per target pixel (scalar stores): used when widths / pitches are not multiples of 4.
// VEC = 4: one thread per 4 consecutive target pixels of a row, word stores; two such groups per iteration with both
// groups' z-buffer loads issued first and both groups' colour gathers second (ncu v1: 44 % issue-active, 48 % of DRAM
// peak, stalled on the key -> gather dependency), 32-bit index arithmetic, every CTA resident.
template <int VEC>
__global__ void __launch_bounds__(kThreads)
    resolve_kernel(unsigned long long *__restrict__ zbuf, const uint8_t *__restrict__ colour, int out_w, int out_h, uint32_t bg_rgb,
                   uint32_t fill_rgb, uint32_t flags, uint8_t *__restrict__ out_rgb, int64_t rgb_pitch, uint8_t *__restrict__ out_mask,
                   int64_t mask_pitch, float *__restrict__ out_depth, int64_t depth_pitch, int32_t *__restrict__ out_ids,
                   const uint8_t *__restrict__ touched, uint8_t *__restrict__ touched_clear) {
    constexpr int G = VEC == 4 ? 2 : 1;  // groups in flight per thread
    const uint32_t groups_per_row = (uint32_t)out_w / VEC;
    const uint32_t n_groups = groups_per_row * (uint32_t)out_h, stride = gridDim.x * kThreads;
    const bool collide = flags & MDVT_FLAG_BG_COLLIDE, reset = flags & MDVT_FLAG_RESET_ZBUF, mask_rgb = flags & MDVT_FLAG_MASK_RGB;
    const uint64_t keep = l2_keep_policy();
    for (uint32_t g0 = blockIdx.x * kThreads + threadIdx.x; g0 < n_groups; g0 += stride * G) {
        unsigned long long key[G][4];
        uint32_t px[G][4], mk[G][4];
        int rows[G], cols[G];
        bool on[G], live[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint32_t gidx = g0 + j * stride;
            on[j] = gidx < n_groups;
            rows[j] = (int)(gidx / groups_per_row);
            cols[j] = (int)(gidx - (uint32_t)rows[j] * groups_per_row) * VEC;
            live[j] = on[j];
            if (on[j] && touched) {
                const uint32_t tt = (uint32_t)rows[j] * (uint32_t)out_w + (uint32_t)cols[j], seg = tt >> kSegShift;
                live[j] = touched[seg] != 0;
                if ((tt & ((1u << kSegShift) - 1)) == 0) touched_clear[seg] = 0;  // the OTHER plane: next frame's
#pragma unroll
                for (int k = 0; k < 4; ++k) key[j][k] = MDVT_ZBUF_EMPTY;
            }
            if (live[j]) {
                const uint32_t t0 = (uint32_t)rows[j] * (uint32_t)out_w + (uint32_t)cols[j];
                if (VEC == 4) {
                    const ulonglong2 a = ld_u64x2_keep(zbuf + t0, keep);
                    const ulonglong2 b = ld_u64x2_keep(zbuf + t0 + 2, keep);
                    key[j][0] = a.x; key[j][1] = a.y; key[j][2] = b.x; key[j][3] = b.y;
                } else {
                    key[j][0] = ld_u64_keep(zbuf + t0, keep);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            if (!on[j]) continue;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                bool hole = key[j][k] == MDVT_ZBUF_EMPTY;
                uint32_t c = fill_rgb;
                if (!hole) {
                    c = gather_rgb(colour, (uint32_t)key[j][k]);
                    if (collide && c == bg_rgb) hole = true;
                    if (hole) c = fill_rgb;
                }
                px[j][k] = c;
                mk[j][k] = hole ? 1u : 0u;
            }
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            if (!on[j]) continue;
            const int row = rows[j], col0 = cols[j];
            const uint32_t t0 = (uint32_t)row * (uint32_t)out_w + (uint32_t)col0;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if (out_depth)
                    out_depth[row * depth_pitch + col0 + k] =
                        (key[j][k] == MDVT_ZBUF_EMPTY) ? 0.0f : __uint_as_float((uint32_t)(key[j][k] >> 32));
                if (out_ids) out_ids[t0 + k] = (key[j][k] == MDVT_ZBUF_EMPTY) ? -1 : (int32_t)(uint32_t)key[j][k];
            }
            if (reset && live[j]) {
                if (VEC == 4) {
                    const ulonglong2 e = make_ulonglong2(MDVT_ZBUF_EMPTY, MDVT_ZBUF_EMPTY);
                    st_u64x2_keep(zbuf + t0, e, keep);
                    st_u64x2_keep(zbuf + t0 + 2, e, keep);
                } else {
                    st_u64_keep(zbuf + t0, MDVT_ZBUF_EMPTY, keep);
                }
            }
            if (out_rgb) {
                uint8_t *o = out_rgb + row * rgb_pitch + (int64_t)col0 * 3;
                if (VEC == 4) {
                    uint32_t *ow = reinterpret_cast<uint32_t *>(o);
                    __stcs(ow, px[j][0] | (px[j][1] << 24));
                    __stcs(ow + 1, (px[j][1] >> 8) | (px[j][2] << 16));
                    __stcs(ow + 2, (px[j][2] >> 16) | (px[j][3] << 8));
                } else {
                    o[0] = (uint8_t)px[j][0]; o[1] = (uint8_t)(px[j][0] >> 8); o[2] = (uint8_t)(px[j][0] >> 16);
                }
            }
            if (out_mask) {
                if (mask_rgb) {
                    uint8_t *o = out_mask + row * mask_pitch + (int64_t)col0 * 3;
                    uint32_t m[4];
#pragma unroll
                    for (int k = 0; k < VEC; ++k) m[k] = mk[j][k] ? bg_rgb : 0u;
                    if (VEC == 4) {
                        uint32_t *ow = reinterpret_cast<uint32_t *>(o);
                        __stcs(ow, m[0] | (m[1] << 24));
                        __stcs(ow + 1, (m[1] >> 8) | (m[2] << 16));
                        __stcs(ow + 2, (m[2] >> 16) | (m[3] << 8));
                    } else {
                        o[0] = (uint8_t)m[0]; o[1] = (uint8_t)(m[0] >> 8); o[2] = (uint8_t)(m[0] >> 16);
                    }
                } else {
                    uint8_t *o = out_mask + row * mask_pitch + col0;
                    if (VEC == 4) {
                        __stcs(reinterpret_cast<uint32_t *>(o),
                               (mk[j][0] * 0xFFu) | ((mk[j][1] * 0xFFu) << 8) | ((mk[j][2] * 0xFFu) << 16) | ((mk[j][3] * 0xFFu) << 24));
                    } else {
                        o[0] = mk[j][0] ? 255 : 0;
                    }
                }
            }
        }
    }
}

// Resident CTAs per SM of a kernel (occupancy API), looked up once per kernel.
template <typename K>
static int resident_ctas(K kernel, int threads) {
    int ctas = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kernel, threads, 0) != cudaSuccess || ctas < 1) ctas = 1;
    return ctas;
}

// K3, row form (out_w % 4 == 0, aligned planes, no id plane): blockIdx.x tiles the 4-pixel groups of a row, blockIdx.y
// strides over the rows two at a time (both rows' z-buffer loads first, then both rows' gathers), so there is no
// division, and a warp whose 128 target pixels lie in untouched segments takes a ~25-instruction path that only
// writes fill colour and mask (ncu v4: the one-kernel-fits-all resolve spent 56 instructions per pixel).
constexpr int kRowResolveThreads = 160;  // 640 px per CTA row chunk: divides 640 / 1280 / 1920 / 3840

struct ViewStrides {  // plane of view v relative to view 0: z-buffer slots, RGB / mask bytes, depth floats
    int64_t zbuf, rgb, mask, depth;
};

template <bool DEPTH>
__global__ void __launch_bounds__(kRowResolveThreads)
    resolve_rows_kernel(unsigned long long *__restrict__ zbuf, const uint8_t *__restrict__ colour, int out_w, int out_h, uint32_t bg_rgb,
                        uint32_t fill_rgb, uint32_t flags, uint8_t *__restrict__ out_rgb, int64_t rgb_pitch, uint8_t *__restrict__ out_mask,
                        int64_t mask_pitch, float *__restrict__ out_depth, int64_t depth_pitch, const uint8_t *__restrict__ touched,
                        uint8_t *__restrict__ touched_clear, ViewStrides vs) {
    const int g = blockIdx.x * kRowResolveThreads + threadIdx.x;
    if (g >= out_w / 4) return;
    if (blockIdx.z) {  // several views in one launch: view v = blockIdx.z works on planes v strides further on
        const int64_t v = blockIdx.z;
        zbuf += v * vs.zbuf;
        if (out_rgb) out_rgb += v * vs.rgb;
        if (out_mask) out_mask += v * vs.mask;
        if (DEPTH) out_depth += v * vs.depth;
    }
    const int col0 = g * 4;
    const bool collide = flags & MDVT_FLAG_BG_COLLIDE, reset = flags & MDVT_FLAG_RESET_ZBUF, mask_rgb = flags & MDVT_FLAG_MASK_RGB;
    const uint64_t keep = l2_keep_policy();
    const int stride = gridDim.y;
    for (int row0 = blockIdx.y; row0 < out_h; row0 += 2 * stride) {
        uint32_t hi[2][4], id[2][4];
        bool on[2], live[2];
        uint32_t t0[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {  // both rows' flags first, then both rows' keys
            const int row = row0 + j * stride;
            on[j] = row < out_h;
            t0[j] = (uint32_t)row * (uint32_t)out_w + (uint32_t)col0;
            live[j] = on[j];
            if (on[j] && touched) live[j] = touched[t0[j] >> kSegShift] != 0;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (on[j] && touched && (t0[j] & ((1u << kSegShift) - 1)) == 0) touched_clear[t0[j] >> kSegShift] = 0;  // the OTHER plane: next frame's
            if (live[j]) {
                const ulonglong2 a = ld_u64x2_keep(zbuf + t0[j], keep);
                const ulonglong2 b = ld_u64x2_keep(zbuf + t0[j] + 2, keep);
                hi[j][0] = (uint32_t)(a.x >> 32); id[j][0] = (uint32_t)a.x;
                hi[j][1] = (uint32_t)(a.y >> 32); id[j][1] = (uint32_t)a.y;
                hi[j][2] = (uint32_t)(b.x >> 32); id[j][2] = (uint32_t)b.x;
                hi[j][3] = (uint32_t)(b.y >> 32); id[j][3] = (uint32_t)b.y;
            }
        }
        // all colour gathers of both rows issued back to back, branch-free (an empty slot reads source pixel 0 and is
        // discarded): one memory latency per iteration instead of up to eight dependent ones
        uint32_t cg[2][4];
        if (live[0] || live[1]) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) cg[j][k] = gather_rgb(colour, (live[j] && hi[j][k] != 0xFFFFFFFFu) ? id[j][k] : 0u);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (!on[j]) continue;
            const int row = row0 + j * stride;
            uint32_t px[4] = {fill_rgb, fill_rgb, fill_rgb, fill_rgb};
            uint32_t holes = 0xF;  // bit k: pixel k is a hole
            if (live[j]) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (hi[j][k] != 0xFFFFFFFFu && !(collide && cg[j][k] == bg_rgb)) {  // a filled slot holds positive finite float bits there
                        px[k] = cg[j][k];
                        holes &= ~(1u << k);
                    }
                }
                if (reset) {
                    const ulonglong2 e = make_ulonglong2(MDVT_ZBUF_EMPTY, MDVT_ZBUF_EMPTY);
                    unsigned long long *z = zbuf + (uint32_t)row * (uint32_t)out_w + (uint32_t)col0;
                    st_u64x2_keep(z, e, keep);
                    st_u64x2_keep(z + 2, e, keep);
                }
            }
            if (DEPTH) {
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live[j]) {
                    d.x = hi[j][0] == 0xFFFFFFFFu ? 0.0f : __uint_as_float(hi[j][0]);
                    d.y = hi[j][1] == 0xFFFFFFFFu ? 0.0f : __uint_as_float(hi[j][1]);
                    d.z = hi[j][2] == 0xFFFFFFFFu ? 0.0f : __uint_as_float(hi[j][2]);
                    d.w = hi[j][3] == 0xFFFFFFFFu ? 0.0f : __uint_as_float(hi[j][3]);
                }
                float *o = out_depth + row * depth_pitch + col0;
                o[0] = d.x; o[1] = d.y; o[2] = d.z; o[3] = d.w;
            }
            if (out_rgb) {
                uint32_t *ow = reinterpret_cast<uint32_t *>(out_rgb + row * rgb_pitch + (int64_t)col0 * 3);
                __stcs(ow, px[0] | (px[1] << 24));
                __stcs(ow + 1, (px[1] >> 8) | (px[2] << 16));
                __stcs(ow + 2, (px[2] >> 16) | (px[3] << 8));
            }
            if (out_mask) {
                if (mask_rgb) {
                    uint32_t m[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) m[k] = (holes >> k) & 1 ? bg_rgb : 0u;
                    uint32_t *ow = reinterpret_cast<uint32_t *>(out_mask + row * mask_pitch + (int64_t)col0 * 3);
                    __stcs(ow, m[0] | (m[1] << 24));
                    __stcs(ow + 1, (m[1] >> 8) | (m[2] << 16));
                    __stcs(ow + 2, (m[2] >> 16) | (m[3] << 8));
                } else {
                    // bits 0..3 -> bytes 0x00 / 0xFF
                    const uint32_t spread = ((holes & 1) | ((holes & 2) << 7) | ((holes & 4) << 14) | ((holes & 8) << 21)) * 0xFFu;
                    __stcs(reinterpret_cast<uint32_t *>(out_mask + row * mask_pitch + col0), spread);
                }
            }
        }
    }
}

// K3 for colour-keyed z-buffers (the frame loops): the low key word already is the winner's colour, so this is a
// streaming pass -- 2 x LDG.128 of keys per 4 target pixels, PRMT packing, word stores -- with no dependent gather.
// Same row form as resolve_rows_kernel (blockIdx.x tiles the 4-pixel groups of a row, blockIdx.y strides over row
// pairs, blockIdx.z = view) and the same per-segment touched flags.  VEC = 1: any width / alignment, one pixel per thread.
template <bool DEPTH, int VEC>
__global__ void __launch_bounds__(kRowResolveThreads)
    resolve_ckey_kernel(unsigned long long *__restrict__ zbuf, int out_w, int out_h, uint32_t bg_rgb, uint32_t fill_rgb, uint32_t flags,
                        uint8_t *__restrict__ out_rgb, int64_t rgb_pitch, uint8_t *__restrict__ out_mask, int64_t mask_pitch,
                        float *__restrict__ out_depth, int64_t depth_pitch, const uint8_t *__restrict__ touched,
                        uint8_t *__restrict__ touched_clear, ViewStrides vs, uint32_t epoch) {
    const int g = blockIdx.x * kRowResolveThreads + threadIdx.x;
    if (g >= (out_w + VEC - 1) / VEC) return;
    if (blockIdx.z) {
        const int64_t v = blockIdx.z;
        zbuf += v * vs.zbuf;
        if (out_rgb) out_rgb += v * vs.rgb;
        if (out_mask) out_mask += v * vs.mask;
        if (DEPTH) out_depth += v * vs.depth;
    }
    const int col0 = g * VEC;
    const bool collide = flags & MDVT_FLAG_BG_COLLIDE, mask_rgb = flags & MDVT_FLAG_MASK_RGB;
    const uint64_t keep = l2_keep_policy();
    const int stride = gridDim.y;
    for (int row0 = blockIdx.y; row0 < out_h; row0 += 2 * stride) {
        uint32_t hi[2][VEC], lo[2][VEC];
        bool on[2], live[2];
        uint32_t t0[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int row = row0 + j * stride;
            on[j] = row < out_h;
            t0[j] = (uint32_t)row * (uint32_t)out_w + (uint32_t)col0;
            live[j] = on[j];
            if (on[j] && touched) live[j] = touched[t0[j] >> kSegShift] != 0;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (on[j] && touched && (t0[j] & ((1u << kSegShift) - 1)) == 0) touched_clear[t0[j] >> kSegShift] = 0;  // the OTHER plane: next frame's
#pragma unroll
            for (int k = 0; k < VEC; ++k) hi[j][k] = 0xFFFFFFFFu, lo[j][k] = 0xFFFFFFFFu;
            if (live[j]) {
                if (VEC == 4) {
                    const ulonglong2 a = ld_u64x2_keep(zbuf + t0[j], keep);
                    const ulonglong2 b = ld_u64x2_keep(zbuf + t0[j] + 2, keep);
                    hi[j][0] = (uint32_t)(a.x >> 32); lo[j][0] = (uint32_t)a.x;
                    hi[j][1] = (uint32_t)(a.y >> 32); lo[j][1] = (uint32_t)a.y;
                    hi[j][2] = (uint32_t)(b.x >> 32); lo[j][2] = (uint32_t)b.x;
                    hi[j][3] = (uint32_t)(b.y >> 32); lo[j][3] = (uint32_t)b.y;
                } else {
                    const unsigned long long a = ld_u64_keep(zbuf + t0[j], keep);
                    hi[j][0] = (uint32_t)(a >> 32); lo[j][0] = (uint32_t)a;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (!on[j]) continue;
            const int row = row0 + j * stride;
            uint32_t px[VEC], holes = 0;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const uint32_t rgb = lo[j][k] & 0xFFFFFFu;
                const bool drawn = (hi[j][k] >> 24) == epoch;   // anything else is an older frame's key or the cleared plane
                const bool hole = !drawn || (collide && rgb == bg_rgb);
                px[k] = hole ? fill_rgb : rgb;
                holes |= hole ? (1u << k) : 0u;
            }
            if (DEPTH) {
                float *o = out_depth + row * depth_pitch + col0;
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    o[k] = (hi[j][k] >> 24) == epoch ? __uint_as_float(((hi[j][k] & 0xFFFFFFu) << 7) | (lo[j][k] >> 25)) : 0.0f;
            }
            if (out_rgb) {
                uint8_t *o = out_rgb + row * rgb_pitch + (int64_t)col0 * 3;
                if (VEC == 4) {
                    uint32_t *ow = reinterpret_cast<uint32_t *>(o);
                    __stcs(ow, px[0] | (px[1] << 24));
                    __stcs(ow + 1, (px[1] >> 8) | (px[2] << 16));
                    __stcs(ow + 2, (px[2] >> 16) | (px[3] << 8));
                } else {
                    o[0] = (uint8_t)px[0]; o[1] = (uint8_t)(px[0] >> 8); o[2] = (uint8_t)(px[0] >> 16);
                }
            }
            if (out_mask) {
                if (mask_rgb) {
                    uint32_t m[VEC];
#pragma unroll
                    for (int k = 0; k < VEC; ++k) m[k] = (holes >> k) & 1 ? bg_rgb : 0u;
                    uint8_t *o = out_mask + row * mask_pitch + (int64_t)col0 * 3;
                    if (VEC == 4) {
                        uint32_t *ow = reinterpret_cast<uint32_t *>(o);
                        __stcs(ow, m[0] | (m[1] << 24));
                        __stcs(ow + 1, (m[1] >> 8) | (m[2] << 16));
                        __stcs(ow + 2, (m[2] >> 16) | (m[3] << 8));
                    } else {
                        o[0] = (uint8_t)m[0]; o[1] = (uint8_t)(m[0] >> 8); o[2] = (uint8_t)(m[0] >> 16);
                    }
                } else if (VEC == 4) {
                    const uint32_t spread = ((holes & 1) | ((holes & 2) << 7) | ((holes & 4) << 14) | ((holes & 8) << 21)) * 0xFFu;
                    __stcs(reinterpret_cast<uint32_t *>(out_mask + row * mask_pitch + col0), spread);
                } else {
                    out_mask[row * mask_pitch + col0] = holes ? 255 : 0;
                }
            }
        }
    }
}

static int grid_for(int64_t work_items) {
    const int64_t blocks = (work_items + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace mdvt

using namespace mdvt;

extern "C" int mdvt_zbuf_clear(uint64_t *zbuf, int64_t n_slots, void *stream) {
    MDVT_REQUIRE(n_slots >= 0, "negative slot count");
    if (n_slots == 0) return MDVT_OK;
    MDVT_REQUIRE(zbuf != nullptr, "zbuf is NULL");
    zbuf_clear_kernel<<<grid_for(n_slots), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<unsigned long long *>(zbuf), n_slots);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

static int launch_project_splat(const void *depth_src, const mdvt_source *src, const ViewPack &pack, const mdvt_view *view_dev, float near_plane,
                                int out_w, int out_h, uint32_t id_offset, unsigned long long *zb, float *out_uvz, cudaStream_t st,
                                uint8_t *touched = nullptr, const uint8_t *colour_key = nullptr, uint32_t epoch = 0);

static int pack_views(const mdvt_view *views_host, int n_views, ViewPack &pack) {
    MDVT_REQUIRE(n_views >= 1 && n_views <= kMaxViews, "n_views must be 1..%d", kMaxViews);
    MDVT_REQUIRE(views_host != nullptr, "views is NULL");
    pack.n = n_views;
    for (int k = 0; k < n_views; ++k) pack.v[k] = views_host[k];
    return MDVT_OK;
}

extern "C" int mdvt_project_splat(const void *depth_src, const mdvt_source *src, const mdvt_view *views_host, int n_views,
                                  float near_plane, int out_w, int out_h, uint32_t id_offset, uint64_t *zbuf, float *out_uvz,
                                  void *stream) {
    if (int rc = check_source(src)) return rc;
    ViewPack pack{};
    if (int rc = pack_views(views_host, n_views, pack)) return rc;
    MDVT_REQUIRE(depth_src && zbuf, "NULL buffer");
    MDVT_REQUIRE(out_w > 0 && out_h > 0, "bad output size %dx%d", out_w, out_h);
    MDVT_REQUIRE((int64_t)src->width * src->height + id_offset <= 0xFFFFFFFFll, "source index does not fit the 32-bit z-buffer payload");
    return launch_project_splat(depth_src, src, pack, nullptr, near_plane, out_w, out_h, id_offset, reinterpret_cast<unsigned long long *>(zbuf),
                                out_uvz, static_cast<cudaStream_t>(stream));
}

extern "C" int mdvt_splat_points(const float *xyz, int64_t n_points, const mdvt_view *views_host, int n_views, float near_plane,
                                 int out_w, int out_h, uint32_t id_offset, uint64_t *zbuf, void *stream) {
    ViewPack pack{};
    if (int rc = pack_views(views_host, n_views, pack)) return rc;
    MDVT_REQUIRE(n_points >= 0, "negative point count");
    MDVT_REQUIRE(out_w > 0 && out_h > 0, "bad output size %dx%d", out_w, out_h);
    if (n_points == 0) return MDVT_OK;
    MDVT_REQUIRE(xyz && zbuf, "NULL buffer");
    MDVT_REQUIRE(n_points + id_offset <= 0xFFFFFFFFll, "point index does not fit the 32-bit z-buffer payload");
    splat_points_kernel<<<grid_for(n_points), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        xyz, n_points, pack, near_plane, out_w, out_h, id_offset, reinterpret_cast<unsigned long long *>(zbuf));
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

static int launch_resolve(unsigned long long *zb, const uint8_t *colour_rgb, int out_w, int out_h, uint32_t bg_rgb, uint32_t fill_rgb,
                          uint32_t flags, uint8_t *out_rgb, int64_t rgb_pitch, uint8_t *out_mask, int64_t mask_pitch, float *out_depth,
                          int64_t depth_pitch, int32_t *out_ids, cudaStream_t st, const uint8_t *touched = nullptr,
                          uint8_t *touched_clear = nullptr, int n_views = 1, ViewStrides vs = ViewStrides{0, 0, 0, 0}) {
    const int mask_bpp = (flags & MDVT_FLAG_MASK_RGB) ? 3 : 1;
    MDVT_REQUIRE(!out_rgb || rgb_pitch >= (int64_t)out_w * 3, "rgb_pitch %lld too small", (long long)rgb_pitch);
    MDVT_REQUIRE(!out_mask || mask_pitch >= (int64_t)out_w * mask_bpp, "mask_pitch %lld too small", (long long)mask_pitch);
    if (depth_pitch == 0) depth_pitch = out_w;
    MDVT_REQUIRE(!out_depth || depth_pitch >= out_w, "depth_pitch %lld too small", (long long)depth_pitch);
    const bool vec4 = (out_w % 4 == 0) && (reinterpret_cast<uintptr_t>(zb) % 16 == 0) &&
                      (!out_rgb || (reinterpret_cast<uintptr_t>(out_rgb) % 4 == 0 && rgb_pitch % 4 == 0)) &&
                      (!out_mask || (reinterpret_cast<uintptr_t>(out_mask) % 4 == 0 && mask_pitch % 4 == 0));
    bg_rgb &= 0xFFFFFF;
    fill_rgb &= 0xFFFFFF;
    MDVT_REQUIRE((int64_t)out_w * out_h < 0x7FFFFFFFll, "target plane must hold fewer than 2^31 pixels");
    static int per_sm4 = 0, per_sm1 = 0;
    if (!per_sm4) per_sm4 = resident_ctas(resolve_kernel<4>, kThreads);
    if (!per_sm1) per_sm1 = resident_ctas(resolve_kernel<1>, kThreads);
    auto grid_of = [&](int64_t items, int per_sm) {
        const int64_t blocks = (items + kThreads - 1) / kThreads, cap = (int64_t)sm_count() * per_sm;
        return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
    };
    if (vec4 && !out_ids && (!out_depth || (reinterpret_cast<uintptr_t>(out_depth) % 4 == 0))) {
        static int per_sm_rows[2] = {0, 0};
        int &per_sm = per_sm_rows[out_depth ? 1 : 0];
        if (!per_sm)
            per_sm = out_depth ? resident_ctas(resolve_rows_kernel<true>, kRowResolveThreads) : resident_ctas(resolve_rows_kernel<false>, kRowResolveThreads);
        const int col_blocks = (out_w / 4 + kRowResolveThreads - 1) / kRowResolveThreads;
        int row_blocks = sm_count() * per_sm / (col_blocks * n_views);
        if (row_blocks < 1) row_blocks = 1;
        if (row_blocks > (out_h + 1) / 2) row_blocks = (out_h + 1) / 2;
        const dim3 grid(col_blocks, row_blocks, n_views);
        if (out_depth)
            resolve_rows_kernel<true><<<grid, kRowResolveThreads, 0, st>>>(zb, colour_rgb, out_w, out_h, bg_rgb, fill_rgb, flags, out_rgb, rgb_pitch,
                                                                          out_mask, mask_pitch, out_depth, depth_pitch, touched, touched_clear, vs);
        else
            resolve_rows_kernel<false><<<grid, kRowResolveThreads, 0, st>>>(zb, colour_rgb, out_w, out_h, bg_rgb, fill_rgb, flags, out_rgb, rgb_pitch,
                                                                           out_mask, mask_pitch, out_depth, depth_pitch, touched, touched_clear, vs);
    } else if (vec4) {
        MDVT_REQUIRE(n_views == 1, "several views per launch need the row form of the resolve");
        resolve_kernel<4><<<grid_of(((int64_t)out_w / 4 * out_h + 1) / 2, per_sm4), kThreads, 0, st>>>(
            zb, colour_rgb, out_w, out_h, bg_rgb, fill_rgb, flags, out_rgb, rgb_pitch, out_mask, mask_pitch, out_depth, depth_pitch, out_ids,
            touched, touched_clear);
    } else {
        MDVT_REQUIRE(n_views == 1, "several views per launch need the row form of the resolve");
        resolve_kernel<1><<<grid_of((int64_t)out_w * out_h, per_sm1), kThreads, 0, st>>>(
            zb, colour_rgb, out_w, out_h, bg_rgb, fill_rgb, flags, out_rgb, rgb_pitch, out_mask, mask_pitch, out_depth, depth_pitch, out_ids,
            touched, touched_clear);
    }
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

// colour-keyed planes -> image / mask / depth, all views of a frame in one launch
static int launch_resolve_ckey(unsigned long long *zb, int out_w, int out_h, uint32_t bg_rgb, uint32_t fill_rgb, uint32_t flags, uint8_t *out_rgb,
                               int64_t rgb_pitch, uint8_t *out_mask, int64_t mask_pitch, float *out_depth, int64_t depth_pitch, cudaStream_t st,
                               const uint8_t *touched, uint8_t *touched_clear, int n_views, ViewStrides vs, uint32_t epoch) {
    const int mask_bpp = (flags & MDVT_FLAG_MASK_RGB) ? 3 : 1;
    MDVT_REQUIRE(!out_rgb || rgb_pitch >= (int64_t)out_w * 3, "rgb_pitch %lld too small", (long long)rgb_pitch);
    MDVT_REQUIRE(!out_mask || mask_pitch >= (int64_t)out_w * mask_bpp, "mask_pitch %lld too small", (long long)mask_pitch);
    if (depth_pitch == 0) depth_pitch = out_w;
    MDVT_REQUIRE(!out_depth || depth_pitch >= out_w, "depth_pitch %lld too small", (long long)depth_pitch);
    MDVT_REQUIRE((int64_t)out_w * out_h < 0x7FFFFFFFll, "target plane must hold fewer than 2^31 pixels");
    auto word_ok = [](const void *p, int64_t a, int64_t b) { return (reinterpret_cast<uintptr_t>(p) % 4 == 0) && a % 4 == 0 && b % 4 == 0; };
    const bool vec4 = out_w % 4 == 0 && reinterpret_cast<uintptr_t>(zb) % 16 == 0 && (vs.zbuf % 2 == 0) &&
                      (!out_rgb || word_ok(out_rgb, rgb_pitch, vs.rgb)) && (!out_mask || word_ok(out_mask, mask_pitch, vs.mask));
    bg_rgb &= 0xFFFFFF;
    fill_rgb &= 0xFFFFFF;
    static int per_sm_of[2][2] = {{0, 0}, {0, 0}};
    int &per_sm = per_sm_of[out_depth ? 1 : 0][vec4 ? 1 : 0];
#define PICK(D, V) resolve_ckey_kernel<D, V>
    if (!per_sm) {
        if (out_depth) per_sm = vec4 ? resident_ctas(PICK(true, 4), kRowResolveThreads) : resident_ctas(PICK(true, 1), kRowResolveThreads);
        else per_sm = vec4 ? resident_ctas(PICK(false, 4), kRowResolveThreads) : resident_ctas(PICK(false, 1), kRowResolveThreads);
    }
    const int vec = vec4 ? 4 : 1;
    const int col_blocks = ((out_w + vec - 1) / vec + kRowResolveThreads - 1) / kRowResolveThreads;
    int row_blocks = sm_count() * per_sm * g_grid_pct() / 100 / (col_blocks * n_views);
    if (row_blocks < 1) row_blocks = 1;
    if (row_blocks > (out_h + 1) / 2) row_blocks = (out_h + 1) / 2;
    const dim3 grid(col_blocks, row_blocks, n_views);
#define GO(D, V)                                                                                                                             \
    PICK(D, V)<<<grid, kRowResolveThreads, 0, st>>>(zb, out_w, out_h, bg_rgb, fill_rgb, flags, out_rgb, rgb_pitch, out_mask, mask_pitch, out_depth, \
                                                    depth_pitch, touched, touched_clear, vs, epoch)
    if (out_depth) { if (vec4) GO(true, 4); else GO(true, 1); }
    else { if (vec4) GO(false, 4); else GO(false, 1); }
#undef GO
#undef PICK
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

static int launch_project_splat(const void *depth_src, const mdvt_source *src, const ViewPack &pack, const mdvt_view *view_dev, float near_plane,
                                int out_w, int out_h, uint32_t id_offset, unsigned long long *zb, float *out_uvz, cudaStream_t st,
                                uint8_t *touched, const uint8_t *colour_key, uint32_t epoch) {
    MDVT_REQUIRE((touched != nullptr) == (view_dev != nullptr), "touched flags go with the device-resident single view");
    MDVT_REQUIRE(out_w < (1 << 22) && out_h < (1 << 22), "target sides must be below 2^22 pixels");
    MDVT_REQUIRE((int64_t)src->width * src->height < 0x7FFFFFFFll && (int64_t)out_w * out_h * pack.n < 0x7FFFFFFFll,
                 "source / target planes must hold fewer than 2^31 pixels");
    SourceCam cam{src->fx, src->fy, src->cx, src->cy, src->grid_sx, src->grid_sy};
    RayPack rays{};
    rays.n = pack.n;
    if (!view_dev)
        for (int k = 0; k < pack.n; ++k)
            rays.v[k] = make_ray_view(cam.fx, cam.fy, cam.cx, cam.cy, cam.sx, cam.sy, pack.v[k].M, pack.v[k].fx, pack.v[k].fy, pack.v[k].cx,
                                      pack.v[k].cy);
    const int col_blocks = (src->width + kSplatThreads - 1) / kSplatThreads;
    MDVT_REQUIRE(!view_dev || colour_key, "the device-resident view goes with colour keys");
    MDVT_REQUIRE(!(colour_key && out_uvz), "no (u, v, z) dump in colour-key mode");
#define CALL(D, B)                                                                                                                   \
    do {                                                                                                                             \
        auto kernel = view_dev ? project_splat_kernel<D, B, true, true>                                                              \
                               : (colour_key ? project_splat_kernel<D, B, false, true> : project_splat_kernel<D, B, false, false>);   \
        static int per_sm_of[3] = {0, 0, 0};                                                                                         \
        int &per_sm = per_sm_of[view_dev ? 2 : (colour_key ? 1 : 0)];                                                                \
        if (!per_sm) per_sm = resident_ctas(kernel, kSplatThreads);                                                                   \
        int row_blocks = sm_count() * per_sm * (colour_key ? g_grid_pct() : 100) / 100 / col_blocks; /* every CTA resident: no second wave */ \
        if (row_blocks < 1) row_blocks = 1;                                                                                          \
        if (row_blocks > src->height) row_blocks = src->height;                                                                      \
        kernel<<<dim3(col_blocks, row_blocks), kSplatThreads, 0, st>>>(depth_src, colour_key, src->width, src->height, src->dec_const, src->depth_scale, cam, \
                                                                     rays, view_dev, near_plane, out_w, out_h, id_offset, zb, out_uvz, touched, epoch << 24); \
    } while (0)
    MDVT_DISPATCH_SOURCE(src->decoder, src->bit16, CALL);
#undef CALL
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_resolve(uint64_t *zbuf, const uint8_t *colour_rgb, int out_w, int out_h, uint32_t bg_rgb, uint32_t fill_rgb,
                            uint32_t flags, uint8_t *out_rgb, int64_t rgb_pitch, uint8_t *out_mask, int64_t mask_pitch,
                            float *out_depth, int64_t depth_pitch, int32_t *out_ids, void *stream) {
    MDVT_REQUIRE(zbuf && colour_rgb, "NULL buffer");
    MDVT_REQUIRE(out_w > 0 && out_h > 0, "bad output size %dx%d", out_w, out_h);
    return launch_resolve(reinterpret_cast<unsigned long long *>(zbuf), colour_rgb, out_w, out_h, bg_rgb, fill_rgb, flags, out_rgb, rgb_pitch,
                          out_mask, mask_pitch, out_depth, depth_pitch, out_ids, static_cast<cudaStream_t>(stream));
}

// An internal second stream (process-wide, one device per process) with its fork / join events: the frame loops below run
// independent work next to the caller's stream on it.
namespace {
constexpr int kAhead = 2, kRing = 4;
struct AuxStream {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, centroid_done[kRing] = {}, frame_done[kRing] = {};
};
int aux_for_current_device(AuxStream **out) {
    static AuxStream aux;
    int dev = 0;
    MDVT_CUDA_TRY(cudaGetDevice(&dev));
    if (aux.device != dev) {
        MDVT_REQUIRE(aux.device == -1, "one process drives one GPU: the library was first used on device %d, now on %d", aux.device, dev);
        MDVT_CUDA_TRY(cudaStreamCreateWithFlags(&aux.stream, cudaStreamNonBlocking));
        MDVT_CUDA_TRY(cudaEventCreateWithFlags(&aux.fork, cudaEventDisableTiming));
        MDVT_CUDA_TRY(cudaEventCreateWithFlags(&aux.join, cudaEventDisableTiming));
        for (int k = 0; k < kRing; ++k) {
            MDVT_CUDA_TRY(cudaEventCreateWithFlags(&aux.centroid_done[k], cudaEventDisableTiming));
            MDVT_CUDA_TRY(cudaEventCreateWithFlags(&aux.frame_done[k], cudaEventDisableTiming));
        }
        aux.device = dev;
    }
    *out = &aux;
    return MDVT_OK;
}
std::mutex g_aux_mutex;  // the internal stream and its events are process-wide: calls from several host threads enqueue one after the other
}  // namespace


// Frame loop of the generic path in one call: per frame K1+K2 for all views (colour-keyed, epoch-tagged), then K3 for all
// views in one launch.  zbuf_sets = 2: `zbuf` holds two sets of n_views planes and the frames alternate between the caller's
// stream (set 0) and the internal second stream (set 1), so that one frame's splat (atomic-rate bound) runs next to the
// other frame's resolve (bandwidth bound).  The planes are cleared once at the end of the call (and every 254 uses of a set).
extern "C" int mdvt_render_views(const void *depth_src, int64_t depth_frame_stride, const uint8_t *colour_rgb, int64_t colour_frame_stride,
                                 int n_frames, const mdvt_source *sources_host, int per_frame_source, const mdvt_view *views_host,
                                 int n_views, float near_plane, int out_w, int out_h, uint64_t *zbuf, int zbuf_sets, uint32_t bg_rgb,
                                 uint32_t fill_rgb, uint32_t flags, const mdvt_plane_layout *rgb_out, const mdvt_plane_layout *mask_out,
                                 const mdvt_plane_layout *depth_out, void *stream) {
    MDVT_REQUIRE(n_frames >= 0, "negative frame count");
    MDVT_REQUIRE(n_views >= 1 && n_views <= kMaxViews, "n_views must be 1..%d", kMaxViews);
    MDVT_REQUIRE(out_w > 0 && out_h > 0, "bad output size %dx%d", out_w, out_h);
    MDVT_REQUIRE(zbuf_sets == 1 || zbuf_sets == 2, "zbuf_sets must be 1 or 2");
    if (n_frames == 0) return MDVT_OK;
    MDVT_REQUIRE(depth_src && colour_rgb && sources_host && views_host && zbuf && rgb_out && rgb_out->base, "NULL buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t out_n = (int64_t)out_w * out_h;
    static const int dbg = getenv("MDVT_DEBUG") ? atoi(getenv("MDVT_DEBUG")) : 0;  // timing aid: 1 no splat, 2 no resolve
    const int sets = n_frames > 1 ? zbuf_sets : 1;
    std::unique_lock<std::mutex> aux_lock(g_aux_mutex, std::defer_lock);
    AuxStream *aux = nullptr;
    cudaStream_t lane[2] = {st, st};
    if (sets == 2) {
        aux_lock.lock();
        if (int rc = aux_for_current_device(&aux)) return rc;
        lane[1] = aux->stream;
        MDVT_CUDA_TRY(cudaEventRecord(aux->fork, st));
        MDVT_CUDA_TRY(cudaStreamWaitEvent(aux->stream, aux->fork, 0));
    }
    for (int f = 0; f < n_frames; ++f) {
        const mdvt_source *src = sources_host + (per_frame_source ? f : 0);
        if (int rc = check_source(src)) return rc;
        MDVT_REQUIRE((int64_t)src->width * src->height <= 0xFFFFFFFFll, "source frame has more than 2^32 pixels");
        ViewPack pack{};
        if (int rc = pack_views(views_host + (int64_t)f * n_views, n_views, pack)) return rc;
        const int set = f % sets, use = f / sets;
        cudaStream_t ls = lane[set];
        unsigned long long *zb = reinterpret_cast<unsigned long long *>(zbuf) + (int64_t)set * n_views * out_n;
        const uint32_t epoch = 254u - (uint32_t)(use % 254);
        const uint8_t *dsrc = static_cast<const uint8_t *>(depth_src) + f * depth_frame_stride;
        const uint8_t *colour_f = colour_rgb + f * colour_frame_stride;
        if (!(dbg & 1))
            if (int rc = launch_project_splat(dsrc, src, pack, nullptr, near_plane, out_w, out_h, 0, zb, nullptr, ls, nullptr, colour_f, epoch)) return rc;
        auto at = [&](const mdvt_plane_layout *L, int v) -> uint8_t * {
            return (L && L->base) ? static_cast<uint8_t *>(L->base) + f * L->frame_stride + v * L->view_stride : nullptr;
        };
        const int64_t mask_pitch = mask_out ? mask_out->row_pitch : 0, depth_pitch = depth_out ? depth_out->row_pitch / 4 : 0;
        // all views in ONE launch (blockIdx.z); the kernel falls back to one pixel per thread for odd widths / alignments
        const ViewStrides vs{out_n, rgb_out->view_stride, mask_out ? mask_out->view_stride : 0, depth_out ? depth_out->view_stride / 4 : 0};
        MDVT_REQUIRE(!at(depth_out, 0) || (depth_out->view_stride % 4 == 0 && depth_out->row_pitch % 4 == 0), "depth planes must be float aligned");
        if (!(dbg & 2))
            if (int rc = launch_resolve_ckey(zb, out_w, out_h, bg_rgb, fill_rgb, flags, at(rgb_out, 0), rgb_out->row_pitch, at(mask_out, 0), mask_pitch,
                                             reinterpret_cast<float *>(at(depth_out, 0)), depth_pitch, ls, nullptr, nullptr, n_views, vs, epoch))
                return rc;
        // the epoch byte is used up (or this was the set's last frame): back to the all-ones plane
        if (use % 254 == 253 || f + sets >= n_frames)
            if (int rc = mdvt_zbuf_clear(reinterpret_cast<uint64_t *>(zb), (int64_t)n_views * out_n, ls)) return rc;
    }
    if (sets == 2) {
        MDVT_CUDA_TRY(cudaEventRecord(aux->join, aux->stream));
        MDVT_CUDA_TRY(cudaStreamWaitEvent(st, aux->join, 0));
    }
    return MDVT_OK;
}

// 3d_view_depthfile.py --render, whole chunk, no host synchronisation: centroid -> device look-at -> splat -> resolve.
// The centroid (+ look-at) of frame f+1 does not depend on frame f, is latency-bound and small, so it runs on an
// internal second stream underneath the splat / resolve of the frames before it (fork / join with events; at most
// kAhead frames ahead so that its input is still in L2 when the splat reads it again).
extern "C" int64_t mdvt_touched_bytes(int out_w, int out_h) {
    if (out_w <= 0 || out_h <= 0) return 0;
    return 2 * (((int64_t)out_h * out_w + (1 << kSegShift) - 1) >> kSegShift);
}

extern "C" int mdvt_novel_view_frames(const void *depth_src, int64_t depth_frame_stride, const uint8_t *colour_rgb, int64_t colour_frame_stride,
                                      int n_frames, const mdvt_source *centroid_src, const mdvt_source *src, const double *K_host,
                                      const double *poses_host, const mdvt_lookat *look, float near_plane, int out_w, int out_h,
                                      uint64_t *zbuf, int zbuf_sets, double *sums_dev, mdvt_view *views_dev, uint8_t *touched, uint32_t bg_rgb,
                                      uint32_t fill_rgb, uint32_t flags, const mdvt_plane_layout *rgb_out, const mdvt_plane_layout *mask_out,
                                      void *stream) {
    MDVT_REQUIRE(n_frames >= 0, "negative frame count");
    MDVT_REQUIRE(out_w > 0 && out_h > 0, "bad output size %dx%d", out_w, out_h);
    MDVT_REQUIRE(zbuf_sets == 1 || zbuf_sets == 2, "zbuf_sets must be 1 or 2");
    if (int rc = check_source(centroid_src)) return rc;
    if (int rc = check_source(src)) return rc;
    if (n_frames == 0) return MDVT_OK;
    MDVT_REQUIRE(depth_src && colour_rgb && K_host && look && zbuf && sums_dev && views_dev && touched && rgb_out && rgb_out->base,
                 "NULL buffer");
    MDVT_REQUIRE(reinterpret_cast<uintptr_t>(views_dev) % 16 == 0, "views_dev must be 16-byte aligned");
    MDVT_REQUIRE(centroid_src->width == src->width && centroid_src->height == src->height, "the two source descriptions differ in size");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Two lanes: even frames on the caller's stream, odd frames on the internal one, each with its own z-buffer plane and
    // pair of touched-flag planes.  A lane runs centroid + look-at -> splat -> resolve of its frames back to back; the
    // latency-bound reduction and the bandwidth-bound resolve of one lane fill in under the atomic-rate-bound splat of the
    // other.
    const int sets = n_frames > 1 ? zbuf_sets : 1;
    std::unique_lock<std::mutex> aux_lock(g_aux_mutex, std::defer_lock);
    AuxStream *aux = nullptr;
    cudaStream_t lane[2] = {st, st};
    ViewPack pack{};
    pack.n = 1;
    const int64_t sums_stride = 4 + MDVT_REDUCE_SCRATCH_DOUBLES;
    const int64_t plane = mdvt_touched_bytes(out_w, out_h) / 2, out_n = (int64_t)out_w * out_h;
    // the "last CTA finishes" counters of the centroid kernels (last scratch double of every frame) and the touched flags:
    // zeroed on the caller's stream before the fork
    MDVT_CUDA_TRY(cudaMemset2DAsync(sums_dev + sums_stride - 1, sums_stride * sizeof(double), 0, sizeof(double), n_frames, st));
    MDVT_CUDA_TRY(cudaMemsetAsync(touched, 0, 2 * plane * sets, st));
    if (sets == 2) {
        aux_lock.lock();
        if (int rc = aux_for_current_device(&aux)) return rc;
        lane[1] = aux->stream;
        MDVT_CUDA_TRY(cudaEventRecord(aux->fork, st));
        MDVT_CUDA_TRY(cudaStreamWaitEvent(aux->stream, aux->fork, 0));
    }
    for (int f = 0; f < n_frames; ++f) {
        const int set = f % sets, use = f / sets;
        cudaStream_t ls = lane[set];
        const uint8_t *dsrc = static_cast<const uint8_t *>(depth_src) + f * depth_frame_stride;
        unsigned long long *zb = reinterpret_cast<unsigned long long *>(zbuf) + (int64_t)set * out_n;
        if (int rc = launch_centroid_lookat(dsrc, centroid_src, K_host, poses_host ? poses_host + 16 * (int64_t)f : nullptr, look,
                                            sums_dev + f * sums_stride, views_dev + f, ls))
            return rc;
        uint8_t *lane_touched = touched + (int64_t)set * 2 * plane;
        uint8_t *cur = lane_touched + (use & 1) * plane, *other = lane_touched + ((use + 1) & 1) * plane;
        const uint32_t epoch = 254u - (uint32_t)(use % 254);
        if (int rc = launch_project_splat(dsrc, src, pack, views_dev + f, near_plane, out_w, out_h, 0, zb, nullptr, ls, cur,
                                          colour_rgb + f * colour_frame_stride, epoch))
            return rc;
        auto at = [&](const mdvt_plane_layout *L) -> uint8_t * {
            return (L && L->base) ? static_cast<uint8_t *>(L->base) + f * L->frame_stride : nullptr;
        };
        if (int rc = launch_resolve_ckey(zb, out_w, out_h, bg_rgb, fill_rgb, flags, at(rgb_out), rgb_out->row_pitch, at(mask_out),
                                         mask_out ? mask_out->row_pitch : 0, nullptr, 0, ls, cur, other, 1, ViewStrides{0, 0, 0, 0}, epoch))
            return rc;
        if (use % 254 == 253 || f + sets >= n_frames)  // epoch byte used up / the lane's last frame: the plane goes back to all ones
            if (int rc = mdvt_zbuf_clear(reinterpret_cast<uint64_t *>(zb), out_n, ls)) return rc;
    }
    if (sets == 2) {
        MDVT_CUDA_TRY(cudaEventRecord(aux->join, aux->stream));
        MDVT_CUDA_TRY(cudaStreamWaitEvent(st, aux->join, 0));
    }
    return MDVT_OK;
}
