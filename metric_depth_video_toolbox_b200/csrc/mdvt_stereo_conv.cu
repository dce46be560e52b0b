// Stereo with a convergence rotation (stereo_rerender.py:704-725,831-836 with --convergence_file, the way movie_2_3D
// drives it): ONE fused kernel per batch of frames, no global z-buffer.
//
// Each eye pose is a rotation about the camera's y axis followed by a shift along x.  A rotation about the optical
// centre moves a pixel's ROW by an amount that does not depend on its depth:
//     v' - cy' = (fy'/fy) * (i - cy) / g(j),     g(j) = M[8] * (j - cx)/fx + M[10]     (= Zv / z)
// so the source pixels that land in target row r are, per source column j, the one or two rows next to
//     i* = cy + (r - cy') * g(j) * fy/fy'.
// A CTA therefore owns one TARGET row (of both eyes) at a time: for every source column it takes the three candidate
// rows around i*, rejects those whose predicted v' is not within 0.51 of r, and runs the exact float32 arithmetic of
// the generic path (mdvt_splat.cu: unproject, 3x4 affine, refined-reciprocal divisions, rintf) on the rest.  A
// candidate whose exact rint(v') equals r goes into a shared-memory z-buffer with a 64-bit atomicMin on
// (float_bits(Zv) << 32 | source row offset << 12 | column) -- the same order as the generic path's
// (Zv, source index), so the result is bit-identical to mdvt_project_splat + mdvt_resolve, which the tests assert.
// Phase B gathers the winners' colours straight from global memory (the few source rows involved are L1/L2 hot),
// packs RGB / mask bytes into a staging row and hands it to a TMA bulk store, like the row-local kernel.
// HBM traffic is the algorithmic 14 B/px; the 33 MB z-buffer planes and their read / re-arm passes are gone.
#include "mdvt_common.cuh"

namespace mdvt {

constexpr int kConvThreads = 256;
constexpr unsigned long long kEmpty64 = MDVT_ZBUF_EMPTY;

struct ConvSmemLayout {
    int zbuf_off, out_off, mask_off, total;
};

__host__ __device__ inline ConvSmemLayout conv_smem_layout(int width, int mask_bpp) {
    ConvSmemLayout L;
    int off = 16;
    L.zbuf_off = off; off += 2 * width * 8;          // per eye: W 64-bit slots
    L.out_off = off;  off += 2 * width * 3;          // left | right RGB row
    off = (off + 15) & ~15;
    L.mask_off = off; off += 2 * width * mask_bpp;
    L.total = (off + 15) & ~15;
    return L;
}

// MASK_MODE: 0 none, 1 u8 {0,255}, 2 u8x3 (bg colour / black)
template <int MASK_MODE>
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
    const int tid = threadIdx.x;
    const uint32_t row_bytes = 3u * width;
    const float u_max = (float)(width - 1);

    for (int k = tid; k < 2 * width; k += kConvThreads) s_z[k] = kEmpty64;
    __syncthreads();

    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int frame = unit / height, r = unit - frame * height;
        const mdvt_conv_frame &fc = frames[frame];
        const float dec_const = fc.dec_const, depth_scale = fc.depth_scale, near_plane = fc.near_plane;
        const float sfx = fc.fx, sfy = fc.fy, scx = fc.cx, scy = fc.cy;
        const float rfx = rcp_refined(sfx), rfy = rcp_refined(sfy);
        const uint8_t *dframe = depth_rgb + (int64_t)frame * height * row_bytes;
        const float fr = (float)r;

        // lowest source row any column of either eye can ask for (g is linear in j: its extremes sit at the borders)
        int i_base = height;
        {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const mdvt_view &vw = fc.view[e];
                const float ratio = vw.fy / sfy;
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    const float xn = ((side ? (float)(width - 1) : 0.0f) - scx) / sfx;
                    const float g = vw.M[8] * xn + vw.M[10];
                    const float ip = scy + (fr - vw.cy) * g / ratio;
                    i_base = min(i_base, (int)floorf(ip) - 3);
                }
            }
            i_base = max(i_base, 0);
        }

        // ---- phase A: candidate source pixels of this target row -> shared-memory z-buffers --------
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const mdvt_view &vw = fc.view[e];
            const float ratio = vw.fy / sfy, m8 = vw.M[8], m10 = vw.M[10];
            unsigned long long *zb = s_z + e * width;
            for (int j = tid; j < width; j += kConvThreads) {
                const float fj = (float)j;
                const float g = m8 * ((fj - scx) / sfx) + m10;      // prediction only: any rounding is covered by the 0.51 window
                const float i_star = scy + (fr - vw.cy) * g / ratio;
                const int i0 = __float2int_rn(i_star);
#pragma unroll
                for (int di = -1; di <= 1; ++di) {
                    const int i = i0 + di;
                    if (i < 0 || i >= height) continue;
                    const float v_pred = vw.cy + ratio * ((float)i - scy) / g;
                    if (!(fabsf(v_pred - fr) <= 0.51f)) continue;
                    // exact path from here on: the arithmetic of splat_pixel() in mdvt_splat.cu, operation for operation
                    const uint8_t *px = dframe + ((int64_t)i * width + j) * 3;
                    const float z = __fmul_rn(depth_of<MDVT_DECODE_D1>(code_of<MDVT_DECODE_D1, true>(px[0], 0u, px[2]), dec_const), depth_scale);
                    const float X = div_rn_by(__fmul_rn(__fsub_rn(fj, scx), z), sfx, rfx);
                    const float Y = div_rn_by(__fmul_rn(__fsub_rn((float)i, scy), z), sfy, rfy);
                    const float Xv = affine_row(vw.M, X, Y, z);
                    const float Yv = affine_row(vw.M + 4, X, Y, z);
                    const float Zv = affine_row(vw.M + 8, X, Y, z);
                    const float rz = rcp_refined(Zv);
                    const float u = __fadd_rn(div_rn_by(__fmul_rn(vw.fx, Xv), Zv, rz), vw.cx);
                    const float v = __fadd_rn(div_rn_by(__fmul_rn(vw.fy, Yv), Zv, rz), vw.cy);
                    const float ur = rintf(u), vr = rintf(v);
                    if (Zv > near_plane && vr == fr && ur >= 0.0f && ur <= u_max) {
                        const unsigned long long key =
                            ((unsigned long long)__float_as_uint(Zv) << 32) | ((uint32_t)(i - i_base) << 12) | (uint32_t)j;
                        atomicMin(&zb[(int)ur], key);
                    }
                }
            }
        }
        if (bulk && tid == 0) bulk_wait_read<0>();  // the previous row's staged output has left shared memory
        __syncthreads();

        // ---- phase B: winners -> colours, hole mask, depth; z-buffers re-armed -------------------------
        const uint8_t *cframe = colour_rgb + (int64_t)frame * height * row_bytes;
        for (int t = tid; t < 2 * width; t += kConvThreads) {
            const unsigned long long key = s_z[t];
            s_z[t] = kEmpty64;
            bool hole = key == kEmpty64;
            uint32_t c = fill_rgb;
            if (!hole) {
                const uint32_t payload = (uint32_t)key;
                const uint8_t *sc = cframe + ((int64_t)(i_base + (int)(payload >> 12)) * width + (payload & 0xFFFu)) * 3;
                c = (uint32_t)__ldg(sc) | ((uint32_t)__ldg(sc + 1) << 8) | ((uint32_t)__ldg(sc + 2) << 16);
                if (collide && c == bg_rgb) {
                    hole = true;
                    c = fill_rgb;
                }
            }
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
#define LAUNCH(M)                                                                                                                   \
    do {                                                                                                                            \
        auto kernel = stereo_conv_rows_kernel<M>;                                                                                   \
        MDVT_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));                          \
        int ctas = 0;                                                                                                               \
        MDVT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kernel, kConvThreads, L.total));                         \
        if (ctas < 1) ctas = 1;                                                                                                     \
        int grid = sm_count() * ctas;                                                                                               \
        if (grid > n_units) grid = n_units;                                                                                         \
        kernel<<<grid, kConvThreads, L.total, st>>>(depth_rgb, colour_rgb, n_units, width, height, frames_dev, bg_rgb & 0xFFFFFF,   \
                                                    fill_rgb & 0xFFFFFF, (flags & MDVT_FLAG_BG_COLLIDE) ? 1 : 0, out_sbs, out_mask, \
                                                    out_depth, bulk);                                                               \
    } while (0)
    if (mode == 0) LAUNCH(0);
    else if (mode == 1) LAUNCH(1);
    else LAUNCH(2);
#undef LAUNCH
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}
