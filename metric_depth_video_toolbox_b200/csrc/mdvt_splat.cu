// Generic novel-view path: K1+K2 (decode -> unproject -> pose -> project -> 64-bit atomicMin splat) and
// K3 (resolve: gather colour by winning source index, hole mask, depth plane, z-buffer reset).
//
// The z-buffer is a u64 plane per view: (float_bits(z') << 32) | source_index.  z' > near > 0, so the
// float bit pattern orders like the value and one unsigned atomicMin implements "nearest wins, ties ->
// lowest source index" deterministically.  1080p stereo = 33 MB, 4K single view = 66 MB: both stay
// resident in the 126 MB L2, so the RED traffic does not reach HBM.
#include "mdvt_common.cuh"

namespace mdvt {

constexpr int kThreads = 256;
constexpr int kMaxViews = 4;
constexpr int kSplatThreads = 128;  // divides 640 / 1280 / 1920 / 3840: no idle tail block per row

struct ViewPack {
    mdvt_view v[kMaxViews];
    int n;
};

// One source pixel through every view.  Divisions are correctly rounded (== the float32 model's `/`): the
// reciprocals of fx, fy are refined once per thread, the one of Zv once per view and shared by u and v.
// All index arithmetic is 32-bit (the entry point checks that source and target planes have < 2^31 pixels).
__device__ __forceinline__ void splat_pixel(uint32_t p, int col, int row, float z, const SourceCam &cam, float rfx, float rfy,
                                            const ViewPack &views, float near_plane, int out_w, uint32_t out_n, float u_max, float v_max,
                                            uint32_t id_offset, uint32_t n, unsigned long long *__restrict__ zbuf, float *__restrict__ out_uvz) {
    const float xg = __fmul_rn(__int2float_rn(col), cam.sx);
    const float yg = __fmul_rn(__int2float_rn(row), cam.sy);
    const float X = div_rn_by(__fmul_rn(__fsub_rn(xg, cam.cx), z), cam.fx, rfx);
    const float Y = div_rn_by(__fmul_rn(__fsub_rn(yg, cam.cy), z), cam.fy, rfy);
#pragma unroll
    for (int k = 0; k < kMaxViews; ++k) {
        if (k < views.n) {
            const mdvt_view &vw = views.v[k];
            const float Xv = affine_row(vw.M, X, Y, z);
            const float Yv = affine_row(vw.M + 4, X, Y, z);
            const float Zv = affine_row(vw.M + 8, X, Y, z);
            float u, v;
            if (out_uvz) {  // parity output: IEEE division also where Zv is zero / negative / tiny
                u = __fadd_rn(__fdiv_rn(__fmul_rn(vw.fx, Xv), Zv), vw.cx);
                v = __fadd_rn(__fdiv_rn(__fmul_rn(vw.fy, Yv), Zv), vw.cy);
                float *o = out_uvz + ((int64_t)k * n + p) * 3;
                o[0] = u; o[1] = v; o[2] = Zv;
            } else {
                const float rz = rcp_refined(Zv);
                u = __fadd_rn(div_rn_by(__fmul_rn(vw.fx, Xv), Zv, rz), vw.cx);
                v = __fadd_rn(div_rn_by(__fmul_rn(vw.fy, Yv), Zv, rz), vw.cy);
            }
            const float ur = rintf(u), vr = rintf(v);  // round half to even, like np.round
            // comparisons are false for NaN, so non-finite projections are culled too
            if (Zv > near_plane && ur >= 0.0f && ur <= u_max && vr >= 0.0f && vr <= v_max) {
                const uint32_t t = (uint32_t)k * out_n + (uint32_t)(int)vr * (uint32_t)out_w + (uint32_t)(int)ur;
                const unsigned long long key = ((unsigned long long)__float_as_uint(Zv) << 32) | (id_offset + p);
                atomicMin(zbuf + t, key);
            }
        }
    }
}

// 2-D launch: blockIdx.x tiles the columns, blockIdx.y strides over the rows, so (col, row) need no division and
// the lanes of a warp cover 32 consecutive source pixels (-> mostly consecutive z-buffer slots: few L2 sectors per
// 64-bit RED).  A 4-pixels-per-thread variant with word loads was measured SLOWER (29 vs 17 us at 1080p x 2 views):
// its lanes hit every fourth slot and each RED touches 4x the sectors -- the atomics, not the loads, matter here.
template <int DECODER, bool BIT16>
__global__ void __launch_bounds__(kSplatThreads)
    project_splat_kernel(const void *__restrict__ rgb, int width, int height, float dec_const, float depth_scale, SourceCam cam,
                         ViewPack views, float near_plane, int out_w, int out_h, uint32_t id_offset,
                         unsigned long long *__restrict__ zbuf, float *__restrict__ out_uvz) {
    const uint32_t out_n = (uint32_t)out_w * (uint32_t)out_h, n = (uint32_t)width * (uint32_t)height;
    const float u_max = (float)(out_w - 1), v_max = (float)(out_h - 1);
    const float rfx = rcp_refined(cam.fx), rfy = rcp_refined(cam.fy);
    const int col = blockIdx.x * kSplatThreads + threadIdx.x;
    if (col >= width) return;
    for (int row = blockIdx.y; row < height; row += gridDim.y) {
        const uint32_t p = (uint32_t)row * (uint32_t)width + (uint32_t)col;
        const float z = __fmul_rn(source_depth<DECODER, BIT16>(rgb, p, dec_const), depth_scale);
        splat_pixel(p, col, row, z, cam, rfx, rfy, views, near_plane, out_w, out_n, u_max, v_max, id_offset, n, zbuf, out_uvz);
    }
}

// Explicit points (N, 3) float32 in the frame space of the views' M: the reference's point painter
// (stereo_rerender.py:746-755,814) and render() of point clouds (background cloud, edge points).
__global__ void __launch_bounds__(kThreads)
    splat_points_kernel(const float *__restrict__ xyz, int64_t n, ViewPack views, float near_plane, int out_w, int out_h, uint32_t id_offset,
                        unsigned long long *__restrict__ zbuf) {
    const int64_t out_n = (int64_t)out_w * out_h;
    const float u_max = (float)(out_w - 1), v_max = (float)(out_h - 1);
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
                    atomicMin(zbuf + (int64_t)k * out_n + t, key);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) zbuf_clear_kernel(unsigned long long *__restrict__ zbuf, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        zbuf[i] = MDVT_ZBUF_EMPTY;
}

__device__ __forceinline__ uint32_t gather_rgb(const uint8_t *__restrict__ colour, uint32_t id) {
    const uint8_t *c = colour + (int64_t)id * 3;
    return (uint32_t)__ldg(c) | ((uint32_t)__ldg(c + 1) << 8) | ((uint32_t)__ldg(c + 2) << 16);
}

// One thread per target pixel (scalar stores): used when widths / pitches are not multiples of 4.
// VEC = 4: one thread per 4 consecutive target pixels of a row, word stores.
template <int VEC>
__global__ void __launch_bounds__(kThreads)
    resolve_kernel(unsigned long long *__restrict__ zbuf, const uint8_t *__restrict__ colour, int out_w, int out_h, uint32_t bg_rgb,
                   uint32_t fill_rgb, uint32_t flags, uint8_t *__restrict__ out_rgb, int64_t rgb_pitch, uint8_t *__restrict__ out_mask,
                   int64_t mask_pitch, float *__restrict__ out_depth, int64_t depth_pitch, int32_t *__restrict__ out_ids) {
    const int groups_per_row = out_w / VEC;
    const int64_t n_groups = (int64_t)groups_per_row * out_h;
    const bool collide = flags & MDVT_FLAG_BG_COLLIDE, reset = flags & MDVT_FLAG_RESET_ZBUF, mask_rgb = flags & MDVT_FLAG_MASK_RGB;
    for (int64_t gidx = blockIdx.x * (int64_t)kThreads + threadIdx.x; gidx < n_groups; gidx += (int64_t)gridDim.x * kThreads) {
        const int row = (int)(gidx / groups_per_row), col0 = (int)(gidx - (int64_t)row * groups_per_row) * VEC;
        const int64_t t0 = (int64_t)row * out_w + col0;
        unsigned long long key[4];
        if (VEC == 4) {
            const ulonglong2 a = reinterpret_cast<const ulonglong2 *>(zbuf + t0)[0];
            const ulonglong2 b = reinterpret_cast<const ulonglong2 *>(zbuf + t0)[1];
            key[0] = a.x; key[1] = a.y; key[2] = b.x; key[3] = b.y;
        } else {
            key[0] = zbuf[t0];
        }
        uint32_t px[4], mk[4];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            bool hole = key[k] == MDVT_ZBUF_EMPTY;
            uint32_t c = fill_rgb;
            if (!hole) {
                c = gather_rgb(colour, (uint32_t)key[k]);
                if (collide && c == bg_rgb) hole = true;
                if (hole) c = fill_rgb;
            }
            px[k] = c;
            mk[k] = hole ? 1u : 0u;
            if (out_depth)
                out_depth[row * depth_pitch + col0 + k] = (key[k] == MDVT_ZBUF_EMPTY) ? 0.0f : __uint_as_float((uint32_t)(key[k] >> 32));
            if (out_ids) out_ids[t0 + k] = (key[k] == MDVT_ZBUF_EMPTY) ? -1 : (int32_t)(uint32_t)key[k];
        }
        if (reset) {
            if (VEC == 4) {
                const ulonglong2 e = make_ulonglong2(MDVT_ZBUF_EMPTY, MDVT_ZBUF_EMPTY);
                reinterpret_cast<ulonglong2 *>(zbuf + t0)[0] = e;
                reinterpret_cast<ulonglong2 *>(zbuf + t0)[1] = e;
            } else {
                zbuf[t0] = MDVT_ZBUF_EMPTY;
            }
        }
        if (out_rgb) {
            uint8_t *o = out_rgb + row * rgb_pitch + (int64_t)col0 * 3;
            if (VEC == 4) {
                uint32_t *ow = reinterpret_cast<uint32_t *>(o);
                ow[0] = px[0] | (px[1] << 24);
                ow[1] = (px[1] >> 8) | (px[2] << 16);
                ow[2] = (px[2] >> 16) | (px[3] << 8);
            } else {
                o[0] = (uint8_t)px[0]; o[1] = (uint8_t)(px[0] >> 8); o[2] = (uint8_t)(px[0] >> 16);
            }
        }
        if (out_mask) {
            if (mask_rgb) {
                uint8_t *o = out_mask + row * mask_pitch + (int64_t)col0 * 3;
                uint32_t m[4];
#pragma unroll
                for (int k = 0; k < VEC; ++k) m[k] = mk[k] ? bg_rgb : 0u;
                if (VEC == 4) {
                    uint32_t *ow = reinterpret_cast<uint32_t *>(o);
                    ow[0] = m[0] | (m[1] << 24);
                    ow[1] = (m[1] >> 8) | (m[2] << 16);
                    ow[2] = (m[2] >> 16) | (m[3] << 8);
                } else {
                    o[0] = (uint8_t)m[0]; o[1] = (uint8_t)(m[0] >> 8); o[2] = (uint8_t)(m[0] >> 16);
                }
            } else {
                uint8_t *o = out_mask + row * mask_pitch + col0;
                if (VEC == 4) {
                    *reinterpret_cast<uint32_t *>(o) =
                        (mk[0] * 0xFFu) | ((mk[1] * 0xFFu) << 8) | ((mk[2] * 0xFFu) << 16) | ((mk[3] * 0xFFu) << 24);
                } else {
                    o[0] = mk[0] ? 255 : 0;
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

static int launch_project_splat(const void *depth_src, const mdvt_source *src, const ViewPack &pack, float near_plane, int out_w, int out_h,
                                uint32_t id_offset, unsigned long long *zb, float *out_uvz, cudaStream_t st);

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
    return launch_project_splat(depth_src, src, pack, near_plane, out_w, out_h, id_offset, reinterpret_cast<unsigned long long *>(zbuf), out_uvz,
                                static_cast<cudaStream_t>(stream));
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
                          int64_t depth_pitch, int32_t *out_ids, cudaStream_t st) {
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
    if (vec4) {
        resolve_kernel<4><<<grid_for((int64_t)out_w / 4 * out_h), kThreads, 0, st>>>(
            zb, colour_rgb, out_w, out_h, bg_rgb, fill_rgb, flags, out_rgb, rgb_pitch, out_mask, mask_pitch, out_depth, depth_pitch, out_ids);
    } else {
        resolve_kernel<1><<<grid_for((int64_t)out_w * out_h), kThreads, 0, st>>>(
            zb, colour_rgb, out_w, out_h, bg_rgb, fill_rgb, flags, out_rgb, rgb_pitch, out_mask, mask_pitch, out_depth, depth_pitch, out_ids);
    }
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

static int launch_project_splat(const void *depth_src, const mdvt_source *src, const ViewPack &pack, float near_plane, int out_w, int out_h,
                                uint32_t id_offset, unsigned long long *zb, float *out_uvz, cudaStream_t st) {
    MDVT_REQUIRE((int64_t)src->width * src->height < 0x7FFFFFFFll && (int64_t)out_w * out_h * pack.n < 0x7FFFFFFFll,
                 "source / target planes must hold fewer than 2^31 pixels");
    SourceCam cam{src->fx, src->fy, src->cx, src->cy, src->grid_sx, src->grid_sy};
    const int col_blocks = (src->width + kSplatThreads - 1) / kSplatThreads;
    int row_blocks = (sm_count() * 16 + col_blocks - 1) / col_blocks;  // ~8 resident CTAs per SM, rows strided
    if (row_blocks > src->height) row_blocks = src->height;
    const dim3 grid(col_blocks, row_blocks);
#define CALL(D, B)                                                                                                                 \
    project_splat_kernel<D, B><<<grid, kSplatThreads, 0, st>>>(depth_src, src->width, src->height, src->dec_const, src->depth_scale, cam, pack, \
                                                          near_plane, out_w, out_h, id_offset, zb, out_uvz)
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

// Frame loop of the generic path in one call: per frame K1+K2 for all views, then K3 per view.
extern "C" int mdvt_render_views(const void *depth_src, int64_t depth_frame_stride, const uint8_t *colour_rgb, int64_t colour_frame_stride,
                                 int n_frames, const mdvt_source *sources_host, int per_frame_source, const mdvt_view *views_host,
                                 int n_views, float near_plane, int out_w, int out_h, uint64_t *zbuf, uint32_t bg_rgb, uint32_t fill_rgb,
                                 uint32_t flags, const mdvt_plane_layout *rgb_out, const mdvt_plane_layout *mask_out,
                                 const mdvt_plane_layout *depth_out, void *stream) {
    MDVT_REQUIRE(n_frames >= 0, "negative frame count");
    MDVT_REQUIRE(n_views >= 1 && n_views <= kMaxViews, "n_views must be 1..%d", kMaxViews);
    MDVT_REQUIRE(out_w > 0 && out_h > 0, "bad output size %dx%d", out_w, out_h);
    if (n_frames == 0) return MDVT_OK;
    MDVT_REQUIRE(depth_src && colour_rgb && sources_host && views_host && zbuf && rgb_out && rgb_out->base, "NULL buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long *zb = reinterpret_cast<unsigned long long *>(zbuf);
    const int64_t out_n = (int64_t)out_w * out_h;
    for (int f = 0; f < n_frames; ++f) {
        const mdvt_source *src = sources_host + (per_frame_source ? f : 0);
        if (int rc = check_source(src)) return rc;
        MDVT_REQUIRE((int64_t)src->width * src->height <= 0xFFFFFFFFll, "source frame has more than 2^32 pixels");
        ViewPack pack{};
        if (int rc = pack_views(views_host + (int64_t)f * n_views, n_views, pack)) return rc;
        const uint8_t *dsrc = static_cast<const uint8_t *>(depth_src) + f * depth_frame_stride;
        if (int rc = launch_project_splat(dsrc, src, pack, near_plane, out_w, out_h, 0, zb, nullptr, st)) return rc;
        for (int v = 0; v < n_views; ++v) {
            auto at = [&](const mdvt_plane_layout *L) -> uint8_t * {
                return (L && L->base) ? static_cast<uint8_t *>(L->base) + f * L->frame_stride + v * L->view_stride : nullptr;
            };
            if (int rc = launch_resolve(zb + v * out_n, colour_rgb + f * colour_frame_stride, out_w, out_h, bg_rgb, fill_rgb,
                                        flags | MDVT_FLAG_RESET_ZBUF, at(rgb_out), rgb_out->row_pitch, at(mask_out),
                                        mask_out ? mask_out->row_pitch : 0, reinterpret_cast<float *>(at(depth_out)),
                                        depth_out ? depth_out->row_pitch / 4 : 0, nullptr, st))
                return rc;
        }
    }
    return MDVT_OK;
}
