// Stereo with a convergence rotation (stereo_rerender.py:704-725,831-836 with --convergence_file, the way movie_2_3D
// drives it): ONE fused kernel per batch of frames, no global z-buffer.
//
// Each eye pose is a rotation about the camera's y axis followed by a shift along x.  In ray form (mdvt_common.cuh)
// that leaves B_u = B_z = T_v = T_z = 0: Zv = z * r_z(j) and v' = r_v(i, j) / r_z(j) -- a pixel's target ROW does not
// depend on its depth.  The source pixels that land in target row r are therefore, per source column j, the one or
// two rows next to
//     i* = (r * r_z(j) - C_v(j)) / B_v.
// A CTA owns one TARGET row (of both eyes) at a time: for every source column and eye it takes the row nearest to i*,
// queues the neighbour when the prediction says it can also round into r, and runs the exact float32 arithmetic of
// the generic path (mdvt_splat.cu: 6 FFMA, refined-reciprocal divisions, magic-number rounding) on the candidates.  A
// candidate whose exact rint(v') equals r goes into a shared-memory z-buffer with a 64-bit atomicMin on
// (float_bits(Zv) << 32 | colour) -- the colour-keyed order of the generic frame loop (nearest Zv, then the smallest
// packed colour), so the result is bit-identical to mdvt_render_views with the same cameras, which the tests assert.
// Phase B needs no gather: it unpacks the keys into RGB / mask bytes in a staging row and hands that to a TMA bulk
// store, like the row-local kernel.
// HBM traffic is the algorithmic 14 B/px; the 33 MB z-buffer planes and their read / re-arm passes are gone.
#include <cstdlib>

#include "mdvt_common.cuh"

namespace mdvt {

constexpr unsigned long long kEmpty64 = MDVT_ZBUF_EMPTY;

struct ConvSmemLayout {
    int ray_off, zbuf_off, out_off, mask_off, queue_off, queue_cap, total;
};

__host__ __device__ inline ConvSmemLayout conv_smem_layout(int width, int mask_bpp) {
    ConvSmemLayout L;
    int off = 16;
    L.ray_off = off;  off += 2 * (int)sizeof(RayView);   // both eyes' coefficients of the current frame
    L.zbuf_off = off; off += 2 * width * 8;          // per eye: W 64-bit slots
    L.out_off = off;  off += 2 * width * 3;          // left | right RGB row
    off = (off + 15) & ~15;
    L.mask_off = off; off += 2 * width * mask_bpp;
    off = (off + 15) & ~15;
    L.queue_cap = (width + 1) / 2;                   // second candidates (typically a few per cent of 2W); overflow is evaluated in place
    L.queue_off = off; off += 4 * L.queue_cap;
    L.total = (off + 15) & ~15;
    return L;
}

__device__ __forceinline__ float rcp_approx(float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return r;
}

// Half-width of the |predicted v' - r| window inside which a candidate row goes through the exact arithmetic.  The
// prediction and the exact float32 path differ by < 2e-3 pixel (a few ulp of a coordinate below 4096), so 0.5 + 0.005
// cannot lose a pixel whose exact rint(v') is r; the narrower the window, the fewer second candidates.
constexpr float kRowWindow = 0.505f;
constexpr float kConvMagic = 12582912.0f;  // 1.5 * 2^23
constexpr int kConvMagicBits = 0x4B400000;

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
    RayView *s_ray = reinterpret_cast<RayView *>(smem + L.ray_off);
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
    int ray_frame = -1;
    for (int unit = blockIdx.x * per_cta; unit < unit_end; ++unit) {
        const int frame = unit / height, r = unit - frame * height;
        const mdvt_conv_frame *fc = frames + frame;
        if (frame != ray_frame) {  // a few times per CTA: both eyes' ray coefficients, float64 -> float32 (make_ray_view)
            __syncthreads();       // nobody still reads the previous frame's coefficients
            if (tid < 2) {
                const mdvt_view *vw = &fc->view[tid];
                float M[12];
#pragma unroll
                for (int k = 0; k < 12; ++k) M[k] = __ldg(&vw->M[k]);
                s_ray[tid] = make_ray_view(__ldg(&fc->fx), __ldg(&fc->fy), __ldg(&fc->cx), __ldg(&fc->cy), 1.0f, 1.0f, M, __ldg(&vw->fx),
                                           __ldg(&vw->fy), __ldg(&vw->cx), __ldg(&vw->cy));
            }
            ray_frame = frame;
            __syncthreads();
        }
        const float dec_const = __ldg(&fc->dec_const), depth_scale = __ldg(&fc->depth_scale), near_plane = __ldg(&fc->near_plane);
        const uint8_t *dframe = depth_rgb + (int64_t)frame * height * row_bytes;
        const uint8_t *cframe = colour_rgb + (int64_t)frame * height * row_bytes;
        const float fr = (float)r;
        // the pose is Ry + x-shift: B_u = B_z = T_v = T_z = 0 exactly (P's second column is (0, fy', 0), its last (fx' m3, 0, 0)),
        // so r_u = C_u(j), r_z = C_z(j) and the FFMAs that would add an exact 0 are left out: same float32 results
        float Au[2], Av[2], Az[2], Cu[2], Cv[2], Cz[2], Bv[2], Tu[2], inv_Bv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            Au[e] = s_ray[e].A[0]; Av[e] = s_ray[e].A[1]; Az[e] = s_ray[e].A[2];
            Cu[e] = s_ray[e].C[0]; Cv[e] = s_ray[e].C[1]; Cz[e] = s_ray[e].C[2];
            Bv[e] = s_ray[e].B[1]; Tu[e] = s_ray[e].T[0];
            inv_Bv[e] = rcp_approx(Bv[e]);
        }

        // lowest source row any column of either eye can ask for (i* is monotone in j: its extremes sit at the borders)
        int i_base = height;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const float fj = side ? (float)(width - 1) : 0.0f;
                const float istar = (fr * (Az[e] * fj + Cz[e]) - (Av[e] * fj + Cv[e])) * inv_Bv[e];
                i_base = min(i_base, (int)floorf(istar) - 3);
            }
        }
        i_base = max(i_base, 0);

        // exact path: the float32 arithmetic of splat_pixel() in mdvt_splat.cu for source pixel (i, j) and eye e
        auto evaluate = [&](int e, int i, float cju, float cjv, float cjz, uint32_t red, uint32_t blue, uint32_t rgb) {
            const float z = __fmul_rn(depth_of<MDVT_DECODE_D1>(code_of<MDVT_DECODE_D1, true>(red, 0u, blue), dec_const), depth_scale);
            const float nu = __fmaf_rn(z, cju, Tu[e]);
            const float nv = __fmaf_rn(z, __fmaf_rn(Bv[e], (float)i, cjv), 0.0f);
            const float Zv = __fmaf_rn(z, cjz, 0.0f);
            const float rz = rcp_refined(Zv);
            const float u = div_rn_by(nu, Zv, rz);
            const float v = div_rn_by(nv, Zv, rz);
            // rint + bounds as in splat_pixel(): the integer sits in the mantissa of u + 1.5 * 2^23; anything out of
            // range, infinite or NaN becomes a huge unsigned value
            const uint32_t ui = (uint32_t)(__float_as_int(__fadd_rn(u, kConvMagic)) - kConvMagicBits);
            const int vi = __float_as_int(__fadd_rn(v, kConvMagic)) - kConvMagicBits;
            if (Zv > near_plane && vi == r && ui < (uint32_t)width) {
                // colour-keyed like the generic frame loop (mdvt_splat.cu, CKEY): nearest Zv, then the smallest packed colour
                atomicMin(&s_z[e * width + ui], ((unsigned long long)__float_as_uint(Zv) << 32) | rgb);
            }
        };

        // ---- phase A: candidate source pixels of this target row -> shared-memory z-buffers --------
        // Per source column and eye the row nearest to i* always goes through the exact arithmetic; its neighbour on the
        // other side of i* can also round into row r when i* sits close to a half (a few per cent of the columns).  Those
        // second candidates are queued in shared memory and evaluated densely afterwards instead of diverging here.
        if (tid == 0) *s_qcount = 0;
        __syncthreads();
        // U columns per thread and pass, both eyes: all 4U byte loads are issued before the first is used
        for (int jb = tid; jb < width; jb += U * kConvThreads) {
            int i0[U][2];
            uint32_t red[U][2], blue[U][2], c0[U][2], c1[U][2], c2[U][2], second[U][2];  // second: queue entry of the neighbour row, 0xFFFFFFFF = none
            float cju[U][2], cjv[U][2], cjz[U][2];
            bool in0[U][2];
            // (1) predictions of all columns and eyes: pure arithmetic, nothing long-latency in between
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int j = jb + k * kConvThreads;
                const float fj = (float)j;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    in0[k][e] = false;
                    second[k][e] = 0xFFFFFFFFu;
                    if (j >= width) continue;
                    cju[k][e] = __fmaf_rn(Au[e], fj, Cu[e]);
                    cjv[k][e] = __fmaf_rn(Av[e], fj, Cv[e]);
                    cjz[k][e] = __fmaf_rn(Az[e], fj, Cz[e]);
                    // prediction (approximate on purpose; the window below absorbs its error)
                    const float g1 = __fmul_rn(cjz[k][e], inv_Bv[e]);                           // d i* / d r  (= 1 / step)
                    const float t = __fmaf_rn(fr, g1, -__fmul_rn(cjv[k][e], inv_Bv[e]));        // i*
                    const float tm = __fadd_rn(t, kConvMagic);
                    i0[k][e] = __float_as_int(tm) - kConvMagicBits;                             // rint(i*)
                    const float step = rcp_approx(g1);                                          // d v' / d i
                    const float v0 = __fmul_rn(__fsub_rn(__fsub_rn(tm, kConvMagic), t), step);  // predicted v' - r of row i0
                    in0[k][e] = (uint32_t)i0[k][e] < (uint32_t)height;
                    if (fabsf(v0) >= __fsub_rn(step, kRowWindow)) {  // the neighbour's predicted v' is within the window of r too
                        const int cand = v0 < 0.0f ? i0[k][e] + 1 : i0[k][e] - 1;
                        if ((uint32_t)cand < (uint32_t)height) second[k][e] = ((uint32_t)e << 31) | ((uint32_t)(cand - i_base) << 12) | (uint32_t)j;
                    }
                }
            }
            // (2) all byte loads back to back
#pragma unroll
            for (int k = 0; k < U; ++k)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (in0[k][e]) {
                        const uint32_t off = ((uint32_t)i0[k][e] * (uint32_t)width + (uint32_t)(jb + k * kConvThreads)) * 3u;
                        red[k][e] = __ldg(dframe + off);
                        blue[k][e] = __ldg(dframe + off + 2);
                        c0[k][e] = __ldg(cframe + off);
                        c1[k][e] = __ldg(cframe + off + 1);
                        c2[k][e] = __ldg(cframe + off + 2);
                    }
            // (3) second candidates -> queue (<= one per column and eye)
#pragma unroll
            for (int k = 0; k < U; ++k)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (second[k][e] != 0xFFFFFFFFu) {
                        const uint32_t slot = atomicAdd(s_qcount, 1u);
                        if (slot < (uint32_t)L.queue_cap) {
                            s_queue[slot] = second[k][e];
                        } else {  // queue full (extreme convergence angles only): the candidate is evaluated here, divergently
                            const int i = i_base + (int)((second[k][e] >> 12) & 0x7FFFFu), j = jb + k * kConvThreads;
                            const uint32_t off = ((uint32_t)i * (uint32_t)width + (uint32_t)j) * 3u;
                            evaluate(e, i, cju[k][e], cjv[k][e], cjz[k][e], __ldg(dframe + off), __ldg(dframe + off + 2),
                                     (uint32_t)__ldg(cframe + off) | ((uint32_t)__ldg(cframe + off + 1) << 8) | ((uint32_t)__ldg(cframe + off + 2) << 16));
                        }
                    }
            // (4) exact arithmetic
#pragma unroll
            for (int k = 0; k < U; ++k)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (in0[k][e])
                        evaluate(e, i0[k][e], cju[k][e], cjv[k][e], cjz[k][e], red[k][e], blue[k][e], c0[k][e] | (c1[k][e] << 8) | (c2[k][e] << 16));
        }
        __syncthreads();
        const int queued = min((int)*s_qcount, L.queue_cap);
        for (int q = tid; q < queued; q += kConvThreads) {
            const uint32_t ent = s_queue[q];
            const int e = (int)(ent >> 31), i = i_base + (int)((ent >> 12) & 0x7FFFFu), j = (int)(ent & 0xFFFu);
            const uint32_t off = ((uint32_t)i * (uint32_t)width + (uint32_t)j) * 3u;
            const float fj = (float)j;
            const float a_u = e ? Au[1] : Au[0], a_v = e ? Av[1] : Av[0], a_z = e ? Az[1] : Az[0];
            const float c_u = e ? Cu[1] : Cu[0], c_v = e ? Cv[1] : Cv[0], c_z = e ? Cz[1] : Cz[0];
            const uint32_t red = __ldg(dframe + off), blue = __ldg(dframe + off + 2);
            const uint32_t rgb = (uint32_t)__ldg(cframe + off) | ((uint32_t)__ldg(cframe + off + 1) << 8) | ((uint32_t)__ldg(cframe + off + 2) << 16);
            if (e) evaluate(1, i, __fmaf_rn(a_u, fj, c_u), __fmaf_rn(a_v, fj, c_v), __fmaf_rn(a_z, fj, c_z), red, blue, rgb);
            else evaluate(0, i, __fmaf_rn(a_u, fj, c_u), __fmaf_rn(a_v, fj, c_v), __fmaf_rn(a_z, fj, c_z), red, blue, rgb);
        }
        if (bulk && tid == 0) bulk_wait_read<0>();  // the previous row's staged output has left shared memory
        __syncthreads();

        // ---- phase B: z-buffers -> colours, hole mask, depth; z-buffers re-armed (the low key word IS the colour) -----
        auto finish_colour = [&](unsigned long long key, bool &hole) -> uint32_t {
            const uint32_t c = (uint32_t)key;  // an occupied slot holds 0x00BBGGRR there, the empty one all ones
            hole = c == 0xFFFFFFFFu || (collide && c == bg_rgb);
            return hole ? fill_rgb : c;
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
                uint32_t c[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) c[k] = finish_colour(key[k], hole[k]);
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
                const uint32_t c = finish_colour(key, hole);
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
    static int conv_u = 0, conv_t = 0;  // development switches: MDVT_CONV_U = 1 | 2 columns in flight per thread (both eyes each),
    if (!conv_u) {                       // MDVT_CONV_THREADS = 128 | 256 | 384 | 512
        const char *e = getenv("MDVT_CONV_U");
        conv_u = (e && atoi(e) == 2) ? 2 : 1;
        const char *t = getenv("MDVT_CONV_THREADS");
        conv_t = t ? atoi(t) : 256;
        if (conv_t != 128 && conv_t != 384 && conv_t != 512) conv_t = 256;
    }
#define LAUNCH(M)                                              \
    do {                                                       \
        if (conv_u == 2) LAUNCH_T(M, 256, 2);                  \
        else if (conv_t == 128) LAUNCH_T(M, 128, 1);           \
        else if (conv_t == 384) LAUNCH_T(M, 384, 1);           \
        else if (conv_t == 512) LAUNCH_T(M, 512, 1);           \
        else LAUNCH_T(M, 256, 1);                              \
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
