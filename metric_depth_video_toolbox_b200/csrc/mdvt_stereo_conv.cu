// Stereo with a convergence rotation (stereo_rerender.py:704-725,831-836 with --convergence_file, the way movie_2_3D
// drives it): ONE fused kernel per batch of frames, no global z-buffer.
//
// Each eye pose is a rotation about the camera's y axis followed by a shift along x.  A rotation about the optical
// centre moves a pixel's ROW by an amount that does not depend on its depth:
//     v' - cy' = (fy'/fy) * (i - cy) / g(j),     g(j) = M[8] * (j - cx)/fx + M[10]     (= Zv / z)
// so the source pixels that land in target row r are, per source column j, the one or two rows next to
//     i* = cy + (r - cy') * g(j) * fy/fy'.
// A CTA therefore owns one TARGET row (of both eyes) at a time: for every source column it takes the three candidate
// rows around i*, rejects those whose predicted v' is not within 0.505 of r, and runs the exact float32 arithmetic of
// the generic path (mdvt_splat.cu: unproject, 3x4 affine, refined-reciprocal divisions, rintf) on the rest.  A
// candidate whose exact rint(v') equals r goes into a shared-memory z-buffer with a 64-bit atomicMin on
// (float_bits(Zv) << 32 | source row offset << 12 | column) -- the same order as the generic path's
// (Zv, source index), so the result is bit-identical to mdvt_project_splat + mdvt_resolve, which the tests assert.
// Phase B gathers the winners' colours straight from global memory (the few source rows involved are L1/L2 hot),
// packs RGB / mask bytes into a staging row and hands it to a TMA bulk store, like the row-local kernel.
// HBM traffic is the algorithmic 14 B/px; the 33 MB z-buffer planes and their read / re-arm passes are gone.
#include <cstdlib>

#include "mdvt_common.cuh"

namespace mdvt {

constexpr unsigned long long kEmpty64 = MDVT_ZBUF_EMPTY;

struct ConvSmemLayout {
    int zbuf_off, out_off, mask_off, queue_off, queue_cap, total;
};

__host__ __device__ inline ConvSmemLayout conv_smem_layout(int width, int mask_bpp) {
    ConvSmemLayout L;
    int off = 16;
    L.zbuf_off = off; off += 2 * width * 8;          // per eye: W 64-bit slots
    L.out_off = off;  off += 2 * width * 3;          // left | right RGB row
    off = (off + 15) & ~15;
    L.mask_off = off; off += 2 * width * mask_bpp;
    off = (off + 15) & ~15;
    L.queue_cap = width;                             // second candidates: at most one per column (typically a few per cent)
    L.queue_off = off; off += 4 * L.queue_cap;
    L.total = (off + 15) & ~15;
    return L;
}

// a / b for the PREDICTION side only (which candidate rows to look at): a couple of ulp off is absorbed by the window.
__device__ __forceinline__ float approx_div(float a, float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return __fmul_rn(a, r);
}

// Half-width of the |predicted v' - r| window inside which a candidate row goes through the exact arithmetic.  The
// prediction and the exact float32 path differ by < 1e-3 pixel (a few ulp of a coordinate below 4096), so 0.5 + 0.005
// cannot lose a pixel whose exact rint(v') is r; the narrower the window, the fewer second candidates.
constexpr float kRowWindow = 0.505f;

// MASK_MODE: 0 none, 1 u8 {0,255}, 2 u8x3 (bg colour / black)
template <int MASK_MODE, int kConvThreads, int U>
__global__ void __launch_bounds__(kConvThreads)
    stereo_conv_rows_kernel(const uint8_t *__restrict__ depth_rgb, const uint8_t *__restrict__ colour_rgb, int n_units, int width, int height,
                            const mdvt_conv_frame *__restrict__ frames, uint32_t bg_rgb, uint32_t fill_rgb, int collide,
                            uint8_t *__restrict__ out_sbs, uint8_t *__restrict__ out_mask, float *__restrict__ out_depth, int bulk) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int mask_bpp = MASK_MODE == 2 ? 3 : 1;
    const ConvSmemLayout L = conv_smem_layout(width, mask_bpp);
    unsigned long long *s_z = reinterpret_cast<unsigned long long *>(smem + L.zbuf_off);  // [2][width]
    uint8_t *s_out = smem + L.out_off;
    uint8_t *s_mask = smem + L.mask_off;
    uint32_t *s_queue = reinterpret_cast<uint32_t *>(smem + L.queue_off);
    uint32_t *s_qcount = reinterpret_cast<uint32_t *>(smem);  // bytes [0, 4)
    const int tid = threadIdx.x;
    const uint32_t row_bytes = 3u * width;
    const bool vec4 = (width % 4 == 0) && (!out_depth || (reinterpret_cast<uintptr_t>(out_depth) & 15) == 0);

    for (int k = tid; k < 2 * width; k += kConvThreads) s_z[k] = kEmpty64;
    __syncthreads();

    // Each CTA takes a CONTIGUOUS block of target rows: successive rows need almost the same few source rows, which then
    // come out of L1 instead of L2.
    const int per_cta = (n_units + gridDim.x - 1) / gridDim.x;
    const int unit_end = min(n_units, (int)(blockIdx.x + 1) * per_cta);
    for (int unit = blockIdx.x * per_cta; unit < unit_end; ++unit) {
        const int frame = unit / height, r = unit - frame * height;
        const mdvt_conv_frame *fc = frames + frame;
        const float dec_const = __ldg(&fc->dec_const), depth_scale = __ldg(&fc->depth_scale), near_plane = __ldg(&fc->near_plane);
        const float sfx = __ldg(&fc->fx), sfy = __ldg(&fc->fy), scx = __ldg(&fc->cx), scy = __ldg(&fc->cy);
        const float rfx = rcp_refined(sfx), rfy = rcp_refined(sfy);
        const uint8_t *dframe = depth_rgb + (int64_t)frame * height * row_bytes;
        const float fr = (float)r;

        // lowest source row any column of either eye can ask for (g is linear in j: its extremes sit at the borders)
        int i_base = height;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float m8 = __ldg(&fc->view[e].M[8]), m10 = __ldg(&fc->view[e].M[10]);
            const float inv_ratio = approx_div(sfy, __ldg(&fc->view[e].fy)), vcy = __ldg(&fc->view[e].cy);
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const float g = m8 * approx_div((side ? (float)(width - 1) : 0.0f) - scx, sfx) + m10;
                i_base = min(i_base, (int)floorf(scy + (fr - vcy) * g * inv_ratio) - 3);
            }
        }
        i_base = max(i_base, 0);

        // ---- phase A: candidate source pixels of this target row -> shared-memory z-buffers --------
        // Per source column the row nearest to i* always goes through the exact arithmetic; its neighbour on the other
        // side of i* can also round into row r when i* sits close to a half (a few per cent of the columns).  Those
        // second candidates are queued in shared memory and evaluated densely afterwards instead of diverging here.
#pragma unroll 1
        for (int e = 0; e < 2; ++e) {
            const mdvt_view *vw = &fc->view[e];
            // the pose is Ry + x-shift: M = [m0 0 m2 m3; 0 1 0 0; m8 0 m10 0].  Dropping the terms that multiply an exact 0
            // or add an exact 0 leaves every finite result as the full 3x4 affine of the generic path computes it.
            const float m0 = __ldg(&vw->M[0]), m2 = __ldg(&vw->M[2]), m3 = __ldg(&vw->M[3]), m8 = __ldg(&vw->M[8]), m10 = __ldg(&vw->M[10]);
            const float vfx = __ldg(&vw->fx), vfy = __ldg(&vw->fy), vcx = __ldg(&vw->cx), vcy = __ldg(&vw->cy);
            const float ratio = approx_div(vfy, sfy), inv_ratio = approx_div(sfy, vfy), inv_sfx = approx_div(1.0f, sfx);
            const float dr = (fr - vcy) * inv_ratio;
            unsigned long long *zb = s_z + e * width;
            // exact path: the float32 arithmetic of splat_pixel() in mdvt_splat.cu for source pixel (i, j)
            auto evaluate = [&](int i, int j, uint32_t red, uint32_t blue) {
                const float z = __fmul_rn(depth_of<MDVT_DECODE_D1>(code_of<MDVT_DECODE_D1, true>(red, 0u, blue), dec_const), depth_scale);
                const float X = div_rn_by(__fmul_rn(__fsub_rn((float)j, scx), z), sfx, rfx);
                const float Y = div_rn_by(__fmul_rn(__fsub_rn((float)i, scy), z), sfy, rfy);
                const float Xv = __fadd_rn(__fadd_rn(__fmul_rn(m0, X), __fmul_rn(m2, z)), m3);
                const float Zv = __fadd_rn(__fmul_rn(m8, X), __fmul_rn(m10, z));
                const float rz = rcp_refined(Zv);
                const float u = __fadd_rn(div_rn_by(__fmul_rn(vfx, Xv), Zv, rz), vcx);
                const float v = __fadd_rn(div_rn_by(__fmul_rn(vfy, Y), Zv, rz), vcy);
                // rint + bounds as in splat_pixel(): the integer sits in the mantissa of u + 1.5 * 2^23; anything out of
                // range, infinite or NaN becomes a huge unsigned value
                const uint32_t ui = (uint32_t)(__float_as_int(__fadd_rn(u, 12582912.0f)) - 0x4B400000);
                const int vi = __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000;
                if (Zv > near_plane && vi == r && ui < (uint32_t)width) {
                    const unsigned long long key =
                        ((unsigned long long)__float_as_uint(Zv) << 32) | ((uint32_t)(i - i_base) << 12) | (uint32_t)j;
                    atomicMin(&zb[ui], key);
                }
            };
            if (tid == 0) *s_qcount = 0;
            __syncthreads();
            // U columns per thread and pass: all 2U byte loads are issued before the first is used
            for (int jb = tid; jb < width; jb += U * kConvThreads) {
                int i0[U], i1[U];   // nearest source row; neighbour row that may also round into r (-1: none)
                uint32_t red[U], blue[U];
                bool in0[U];
                // (1) predictions of all U columns: pure arithmetic, nothing long-latency in between
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int j = jb + k * kConvThreads;
                    in0[k] = false;
                    i1[k] = -1;
                    if (j >= width) continue;
                    // prediction (approximate on purpose; the window below absorbs its error): g = Zv / z of this column
                    const float g = __fmaf_rn(m8, __fmul_rn(__fsub_rn((float)j, scx), inv_sfx), m10);
                    i0[k] = __float2int_rn(__fmaf_rn(dr, g, scy));
                    float inv_g;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_g) : "f"(g));
                    const float step = __fmul_rn(ratio, inv_g);                                            // d v' / d i
                    const float v0 = __fsub_rn(__fmaf_rn(__fsub_rn((float)i0[k], scy), step, vcy), fr);    // predicted v' - r of row i0
                    in0[k] = (uint32_t)i0[k] < (uint32_t)height;
                    if (fabsf(v0) >= __fsub_rn(step, kRowWindow)) {  // the neighbour's predicted v' is within the window of r too
                        const int cand = v0 < 0.0f ? i0[k] + 1 : i0[k] - 1;
                        if ((uint32_t)cand < (uint32_t)height) i1[k] = cand;
                    }
                }
                // (2) all 2U byte loads back to back (v3 interleaved them with the queue's warp-aggregated atomics and ncu
                //     showed every column's prediction waiting on the previous column's loads: 27 % of the stall samples)
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    if (in0[k]) {
                        const uint8_t *px = dframe + ((uint32_t)i0[k] * (uint32_t)width + (uint32_t)(jb + k * kConvThreads)) * 3u;
                        red[k] = __ldg(px);
                        blue[k] = __ldg(px + 2);
                    }
                }
                // (3) second candidates -> queue (<= one per column)
#pragma unroll
                for (int k = 0; k < U; ++k)
                    if (i1[k] >= 0) s_queue[atomicAdd(s_qcount, 1u)] = ((uint32_t)(i1[k] - i_base) << 12) | (uint32_t)(jb + k * kConvThreads);
                // (4) exact arithmetic
#pragma unroll
                for (int k = 0; k < U; ++k)
                    if (in0[k]) evaluate(i0[k], jb + k * kConvThreads, red[k], blue[k]);
            }
            __syncthreads();
            const int queued = min((int)*s_qcount, L.queue_cap);
            for (int q = tid; q < queued; q += kConvThreads) {
                const uint32_t ent = s_queue[q];
                const int i = i_base + (int)(ent >> 12), j = (int)(ent & 0xFFFu);
                const uint32_t off = ((uint32_t)i * (uint32_t)width + (uint32_t)j) * 3u;
                evaluate(i, j, __ldg(dframe + off), __ldg(dframe + off + 2));
            }
            __syncthreads();  // the queue is reused by the other eye
        }
        if (bulk && tid == 0) bulk_wait_read<0>();  // the previous row's staged output has left shared memory
        __syncthreads();

        // ---- phase B: winners -> colours, hole mask, depth; z-buffers re-armed -------------------------
        const uint8_t *cframe = colour_rgb + (int64_t)frame * height * row_bytes;
        // byte offset of the winner's colour inside the frame (holes read pixel 0: a harmless, always valid address)
        auto colour_offset = [&](unsigned long long key) -> uint32_t {
            const uint32_t payload = (uint32_t)key;
            // an occupied slot never has an all-ones payload (column <= 4095, row offset < 2^20): one 32-bit test
            return payload == 0xFFFFFFFFu ? 0u : ((uint32_t)(i_base + (int)(payload >> 12)) * (uint32_t)width + (payload & 0xFFFu)) * 3u;
        };
        auto finish_colour = [&](unsigned long long key, uint32_t c, bool &hole) -> uint32_t {
            hole = (uint32_t)key == 0xFFFFFFFFu || (collide && c == bg_rgb);
            return hole ? fill_rgb : c;
        };
        auto colour_of = [&](unsigned long long key, bool &hole) -> uint32_t {
            const uint8_t *sc = cframe + colour_offset(key);
            return finish_colour(key, (uint32_t)__ldg(sc) | ((uint32_t)__ldg(sc + 1) << 8) | ((uint32_t)__ldg(sc + 2) << 16), hole);
        };
        if (vec4) {  // 4 consecutive target pixels per thread: 2 x LDS.128 keys, word stores into the staging rows
            for (int g4 = tid; g4 < 2 * width / 4; g4 += kConvThreads) {
                ulonglong2 *zq = reinterpret_cast<ulonglong2 *>(s_z + 4 * g4);
                const ulonglong2 ka = zq[0], kb = zq[1];
                const ulonglong2 empty2 = make_ulonglong2(kEmpty64, kEmpty64);
                zq[0] = empty2;
                zq[1] = empty2;
                const unsigned long long key[4] = {ka.x, ka.y, kb.x, kb.y};
                bool hole[4];
                uint32_t c[4], b0[4], b1[4], b2[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // all twelve byte loads in flight before any is used
                    const uint8_t *sc = cframe + colour_offset(key[k]);
                    b0[k] = __ldg(sc); b1[k] = __ldg(sc + 1); b2[k] = __ldg(sc + 2);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) c[k] = finish_colour(key[k], b0[k] | (b1[k] << 8) | (b2[k] << 16), hole[k]);
                uint32_t *ow = reinterpret_cast<uint32_t *>(s_out) + 3 * g4;
                ow[0] = c[0] | (c[1] << 24);
                ow[1] = (c[1] >> 8) | (c[2] << 16);
                ow[2] = (c[2] >> 16) | (c[3] << 8);
                if (MASK_MODE == 1) {
                    reinterpret_cast<uint32_t *>(s_mask)[g4] =
                        (hole[0] ? 0xFFu : 0u) | (hole[1] ? 0xFF00u : 0u) | (hole[2] ? 0xFF0000u : 0u) | (hole[3] ? 0xFF000000u : 0u);
                } else if (MASK_MODE == 2) {
                    uint32_t m[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) m[k] = hole[k] ? bg_rgb : 0u;
                    uint32_t *mw = reinterpret_cast<uint32_t *>(s_mask) + 3 * g4;
                    mw[0] = m[0] | (m[1] << 24);
                    mw[1] = (m[1] >> 8) | (m[2] << 16);
                    mw[2] = (m[2] >> 16) | (m[3] << 8);
                }
                if (out_depth) {
                    float4 d;
                    d.x = (uint32_t)key[0] == 0xFFFFFFFFu ? 0.0f : __uint_as_float((uint32_t)(key[0] >> 32));
                    d.y = (uint32_t)key[1] == 0xFFFFFFFFu ? 0.0f : __uint_as_float((uint32_t)(key[1] >> 32));
                    d.z = (uint32_t)key[2] == 0xFFFFFFFFu ? 0.0f : __uint_as_float((uint32_t)(key[2] >> 32));
                    d.w = (uint32_t)key[3] == 0xFFFFFFFFu ? 0.0f : __uint_as_float((uint32_t)(key[3] >> 32));
                    reinterpret_cast<float4 *>(out_depth + (int64_t)unit * 2 * width)[g4] = d;
                }
            }
        } else {
            for (int t = tid; t < 2 * width; t += kConvThreads) {
                const unsigned long long key = s_z[t];
                s_z[t] = kEmpty64;
                bool hole;
                const uint32_t c = colour_of(key, hole);
                uint8_t *o = s_out + 3 * t;
                o[0] = (uint8_t)c; o[1] = (uint8_t)(c >> 8); o[2] = (uint8_t)(c >> 16);
                if (MASK_MODE == 1) {
                    s_mask[t] = hole ? 255 : 0;
                } else if (MASK_MODE == 2) {
                    const uint32_t m = hole ? bg_rgb : 0u;
                    uint8_t *mo = s_mask + 3 * t;
                    mo[0] = (uint8_t)m; mo[1] = (uint8_t)(m >> 8); mo[2] = (uint8_t)(m >> 16);
                }
                if (out_depth) out_depth[(int64_t)unit * 2 * width + t] = key == kEmpty64 ? 0.0f : __uint_as_float((uint32_t)(key >> 32));
            }
        }

        // ---- staged row -> HBM ---------------------------------------------------------------------
        uint8_t *g_out = out_sbs + (int64_t)unit * 2 * row_bytes;
        uint8_t *g_mask = MASK_MODE != 0 ? out_mask + (int64_t)unit * 2 * width * mask_bpp : nullptr;
        if (bulk) {
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                bulk_store(g_out, s_out, 2 * row_bytes);
                if (MASK_MODE != 0) bulk_store(g_mask, s_mask, 2 * width * mask_bpp);
                bulk_commit();
            }
        } else {
            __syncthreads();
            for (uint32_t k = tid; k < 2 * row_bytes; k += kConvThreads) g_out[k] = s_out[k];
            if (MASK_MODE != 0)
                for (int k = tid; k < 2 * width * mask_bpp; k += kConvThreads) g_mask[k] = s_mask[k];
            __syncthreads();
        }
    }
    if (bulk && tid == 0) bulk_wait_all<0>();
}

}  // namespace mdvt

using namespace mdvt;

extern "C" int mdvt_stereo_conv_rows(const uint8_t *depth_rgb, const uint8_t *colour_rgb, int n_frames, int width, int height,
                                     const mdvt_conv_frame *frames_dev, uint32_t bg_rgb, uint32_t fill_rgb, uint32_t flags, uint8_t *out_sbs,
                                     uint8_t *out_mask, float *out_depth, void *stream) {
    MDVT_REQUIRE(n_frames >= 0, "negative frame count");
    MDVT_REQUIRE(width > 0 && height > 0, "bad frame size %dx%d", width, height);
    if (width > 4096) {
        set_error("mdvt_stereo_conv_rows packs the source column into 12 bits: width %d > 4096", width);
        return MDVT_ERR_UNSUPPORTED;
    }
    if (n_frames == 0) return MDVT_OK;
    MDVT_REQUIRE(depth_rgb && colour_rgb && frames_dev && out_sbs, "NULL buffer");
    MDVT_REQUIRE((int64_t)n_frames * height <= 0x7FFFFFFFll, "too many rows in one batch");
    MDVT_REQUIRE((int64_t)width * height * 3 <= 0xFFFFFFFFll, "frame too large for 32-bit byte offsets");
    const int mode = !out_mask ? 0 : ((flags & MDVT_FLAG_MASK_RGB) ? 2 : 1);
    const ConvSmemLayout L = conv_smem_layout(width, mode == 2 ? 3 : 1);
    int dev = 0, smem_optin = 0;
    MDVT_CUDA_TRY(cudaGetDevice(&dev));
    MDVT_CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (L.total > smem_optin) {
        set_error("row of width %d needs %d bytes of shared memory, device offers %d", width, L.total, smem_optin);
        return MDVT_ERR_UNSUPPORTED;
    }
    auto aligned16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const int bulk = (width % 16 == 0) && aligned16(out_sbs) && (!out_mask || aligned16(out_mask));
    const int n_units = n_frames * height;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // 256 threads per CTA (4 CTAs/SM at 1080p): measured 30.5 us/frame; 128 -> 44.9, 160 -> 39.0, 320 -> 31.9.  Unlike the
    // row-local kernel this one is bound by instruction issue and global-load latency, so it wants the warps.
    static int conv_u = 0;  // columns in flight per thread: MDVT_CONV_U = 4 | 8 (development switch)
    if (!conv_u) {
        const char *e = getenv("MDVT_CONV_U");
        conv_u = (e && atoi(e) == 8) ? 8 : 4;  // measured equal (30.6 vs 30.7 us per 1080p frame): loads in flight are not the limiter
    }
#define LAUNCH(M)                     \
    do {                              \
        if (conv_u == 4) LAUNCH_T(M, 256, 4); \
        else LAUNCH_T(M, 256, 8);     \
    } while (0)
#define LAUNCH_T(M, kConvThreads, UU)                                                                                                   \
    do {                                                                                                                            \
        auto kernel = stereo_conv_rows_kernel<M, kConvThreads, UU>;                                                                                   \
        MDVT_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));                          \
        int ctas = 0;                                                                                                               \
        MDVT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kernel, kConvThreads, L.total));                         \
        if (ctas < 1) ctas = 1;                                                                                                     \
        int grid = sm_count() * ctas;                                                                                               \
        if (grid > n_units) grid = n_units;                                                                                         \
        grid = (n_units + ((n_units + grid - 1) / grid) - 1) / ((n_units + grid - 1) / grid); /* no empty CTAs */                                                                                         \
        kernel<<<grid, kConvThreads, L.total, st>>>(depth_rgb, colour_rgb, n_units, width, height, frames_dev, bg_rgb & 0xFFFFFF,   \
                                                    fill_rgb & 0xFFFFFF, (flags & MDVT_FLAG_BG_COLLIDE) ? 1 : 0, out_sbs, out_mask, \
                                                    out_depth, bulk);                                                               \
    } while (0)
    if (mode == 0) LAUNCH(0);
    else if (mode == 1) LAUNCH(1);
    else LAUNCH(2);
#undef LAUNCH
#undef LAUNCH_T
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}
