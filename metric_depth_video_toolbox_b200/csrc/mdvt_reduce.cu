// Per-frame reductions that would otherwise force a full-frame host round trip:
//   * vertex centroid of the unprojected frame (Open3D get_center(), 3d_view_depthfile.py:231 -- the
//     look-at target of the novel-view camera)
//   * (masked) mean depth (find_convergence_depth.py:56-80 -- the convergence depth list)
// Both are float64 sums in a FIXED order (per-thread grid-stride partial -> warp shuffle tree -> block
// tree -> one partial per CTA -> a single-CTA finishing pass), so results are reproducible run to run.
#include "mdvt_common.cuh"

namespace mdvt {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = MDVT_REDUCE_SCRATCH_DOUBLES / 4;

struct Pose12d {
    double m[12];
    int on;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    return v;
}

// Sum 4 doubles over the CTA in a fixed order; the result is valid in thread 0.
__device__ __forceinline__ void block_sum4(double (&a)[4]) {
    __shared__ double s[4][kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k] = warp_sum(a[k]);
        if (lane == 0) s[k][warp] = a[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double v = lane < kThreads / 32 ? s[k][lane] : 0.0;
            a[k] = warp_sum(v);
        }
    }
}

// Vertex sums of one frame.  The per-vertex formula is X = (xg - cx) * z / fx (depth_map_tools.py:1127-1128); its
// SUM over the frame is rearranged so that the per-pixel work is three float64 multiply-adds and no division:
//     sum X = (sum(xg * z) - cx * sum(z)) / fx,   sum Y likewise,   sum Z = sum(z)
// (the pose, being affine, is applied to the mean on the host side of the ABI: mean(T p) = T mean(p)).  The result
// differs from a sum of individually rounded vertices by O(1e-16) relative -- the reference's own mean is a pairwise
// float64 sum, i.e. no bit-exact target exists; tests hold it to 1e-12 relative.
// partial[block] = {sum(xg*z), sum(yg*z), sum(z), count}.  VEC4: 4 pixels per thread from three aligned words.
template <int DECODER, bool BIT16, bool VEC4>
__global__ void __launch_bounds__(kThreads)
    centroid_partial_kernel(const void *__restrict__ src, int width, int64_t n, float dec_const, float depth_scale, float sx, float sy,
                            int stretched, double *__restrict__ partial) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    auto add = [&](int64_t p, float zf) {
        const int row = (int)(p / width), col = (int)(p - (int64_t)row * width);
        const double xg = stretched ? (double)__fmul_rn(__int2float_rn(col), sx) : (double)col;
        const double yg = stretched ? (double)__fmul_rn(__int2float_rn(row), sy) : (double)row;
        const double z = (double)zf;
        acc[0] = fma(xg, z, acc[0]);
        acc[1] = fma(yg, z, acc[1]);
        acc[2] += z;
        acc[3] += 1.0;
    };
    if (VEC4 && DECODER != MDVT_SOURCE_F32) {
        const uint32_t *words = reinterpret_cast<const uint32_t *>(src);
        const int64_t n4 = n / 4;
        for (int64_t g = blockIdx.x * (int64_t)kThreads + threadIdx.x; g < n4; g += (int64_t)gridDim.x * kThreads) {
            const uint32_t w0 = __ldg(words + 3 * g), w1 = __ldg(words + 3 * g + 1), w2 = __ldg(words + 3 * g + 2);
            const uint32_t r[4] = {w0 & 0xFF, w0 >> 24, (w1 >> 16) & 0xFF, (w2 >> 8) & 0xFF};
            const uint32_t gr[4] = {(w0 >> 8) & 0xFF, w1 & 0xFF, w1 >> 24, (w2 >> 16) & 0xFF};
            const uint32_t b[4] = {(w0 >> 16) & 0xFF, (w1 >> 8) & 0xFF, w2 & 0xFF, w2 >> 24};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                add(4 * g + k, __fmul_rn(depth_of<DECODER>(code_of<DECODER, BIT16>(r[k], gr[k], b[k]), dec_const), depth_scale));
        }
    } else {
        for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads)
            add(p, __fmul_rn(source_depth<DECODER, BIT16>(src, p, dec_const), depth_scale));
    }
    block_sum4(acc);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) partial[blockIdx.x * 4 + k] = acc[k];
    }
}

// Single CTA: fixed-order sum of the partials, then the closed form above and the optional pose.
__global__ void __launch_bounds__(kThreads) finish_centroid_kernel(const double *__restrict__ partial, int n_blocks, double fx, double fy,
                                                                   double cx, double cy, Pose12d pose, double *__restrict__ out) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < n_blocks; b += kThreads) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += partial[b * 4 + k];
    }
    block_sum4(acc);
    if (threadIdx.x == 0) {
        const double sz = acc[2], cnt = acc[3];
        double X = (acc[0] - cx * sz) / fx, Y = (acc[1] - cy * sz) / fy, Z = sz;
        if (pose.on) {  // sum(T p) = T[:, :3] sum(p) + count * T[:, 3]
            const double *m = pose.m;
            const double x2 = m[0] * X + m[1] * Y + m[2] * Z + m[3] * cnt;
            const double y2 = m[4] * X + m[5] * Y + m[6] * Z + m[7] * cnt;
            const double z2 = m[8] * X + m[9] * Y + m[10] * Z + m[11] * cnt;
            X = x2; Y = y2; Z = z2;
        }
        out[0] = X; out[1] = Y; out[2] = Z; out[3] = cnt;
    }
}

template <int DECODER, bool BIT16>
__global__ void __launch_bounds__(kThreads)
    depth_sum_partial_kernel(const void *__restrict__ src, int64_t n, float dec_const, const uint8_t *__restrict__ mask, int mask_gt,
                             double *__restrict__ partial) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};  // sum, count, sum of squares, unused
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        if (mask && (int)__ldg(mask + p) <= mask_gt) continue;
        const double d = (double)source_depth<DECODER, BIT16>(src, p, dec_const);
        acc[0] += d; acc[1] += 1.0; acc[2] += d * d;
    }
    block_sum4(acc);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) partial[blockIdx.x * 4 + k] = acc[k];
    }
}

__global__ void __launch_bounds__(kThreads) finish_sum4_kernel(const double *__restrict__ partial, int n_blocks, double *__restrict__ out) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < n_blocks; b += kThreads) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += partial[b * 4 + k];
    }
    block_sum4(acc);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) out[k] = acc[k];
    }
}

static int reduce_grid(int64_t n) {
    const int64_t blocks = (n + kThreads * 8 - 1) / (kThreads * 8);  // >= 8 elements per thread
    const int64_t cap = sm_count() * 4 < kMaxBlocks ? sm_count() * 4 : kMaxBlocks;
    return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace mdvt

using namespace mdvt;

extern "C" int mdvt_centroid(const void *depth_src, const mdvt_source *src, const double *K_host, const double *pose_host,
                             double *out_sums, void *stream) {
    if (int rc = check_source(src)) return rc;
    MDVT_REQUIRE(depth_src && K_host && out_sums, "NULL buffer");
    const int64_t n = (int64_t)src->width * src->height;
    Pose12d pose{};
    if (pose_host) {
        for (int k = 0; k < 12; ++k) pose.m[k] = pose_host[k];
        pose.on = 1;
    }
    const int stretched = !(src->grid_sx == 1.0f && src->grid_sy == 1.0f);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (n % 4 == 0) && (reinterpret_cast<uintptr_t>(depth_src) % 4 == 0);
    const int grid = reduce_grid(vec4 ? n / 4 : n);
    double *partial = out_sums + 4;
#define CALL(D, B)                                                                                                            \
    do {                                                                                                                      \
        if (vec4)                                                                                                             \
            centroid_partial_kernel<D, B, true><<<grid, kThreads, 0, st>>>(depth_src, src->width, n, src->dec_const, src->depth_scale, \
                                                                           src->grid_sx, src->grid_sy, stretched, partial);   \
        else                                                                                                                  \
            centroid_partial_kernel<D, B, false><<<grid, kThreads, 0, st>>>(depth_src, src->width, n, src->dec_const, src->depth_scale, \
                                                                            src->grid_sx, src->grid_sy, stretched, partial);  \
    } while (0)
    MDVT_DISPATCH_SOURCE(src->decoder, src->bit16, CALL);
#undef CALL
    MDVT_CUDA_TRY(cudaGetLastError());
    finish_centroid_kernel<<<1, kThreads, 0, st>>>(partial, grid, K_host[0], K_host[1], K_host[2], K_host[3], pose, out_sums);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_depth_sum(const void *depth_src, int64_t n_pixels, int decoder, int bit16, float dec_const, const uint8_t *mask,
                              int mask_gt, double *out_sums, void *stream) {
    MDVT_REQUIRE(n_pixels >= 0, "negative pixel count");
    if (int rc = check_decoder(decoder, bit16, true)) return rc;
    MDVT_REQUIRE(out_sums != nullptr, "out_sums is NULL");
    MDVT_REQUIRE(n_pixels == 0 || depth_src != nullptr, "depth_src is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = reduce_grid(n_pixels);
    double *partial = out_sums + 4;
#define CALL(D, B) depth_sum_partial_kernel<D, B><<<grid, kThreads, 0, st>>>(depth_src, n_pixels, dec_const, mask, mask_gt, partial)
    MDVT_DISPATCH_SOURCE(decoder, bit16, CALL);
#undef CALL
    MDVT_CUDA_TRY(cudaGetLastError());
    finish_sum4_kernel<<<1, kThreads, 0, st>>>(partial, grid, out_sums);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}
