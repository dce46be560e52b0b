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

// finish_centroid_kernel + the camera of 3d_view_depthfile.py:231-241 on the device: look-at point = centroid with the
// --tx/--ty/--tz overrides, cam_look_at (depth_map_tools.py:1618-1638: r, u, f as columns, translation (px, py, -pz)),
// render()'s Y scale (:1528-1552) and the frame's pose folded into one 3x4, rounded to float32 into `view` -- the
// per-frame host round trip (reduce -> D2H -> NumPy look-at -> launch) of v1 is gone; float64, one rounding per written
// operation, evaluation order of the NumPy helpers (a 1-ulp float64 difference against BLAS-evaluated dot products can
// survive the float32 rounding of an entry only with probability ~1e-8).
struct Pose16d {
    double m[16];
    int on;
};

// Runs in the LAST CTA of centroid_partial_kernel to finish (threadfence + counter): no second launch, and the partials
// are still in L2.  The sum over partials keeps its fixed order (by CTA index), so the result is reproducible.
struct LookAtFinish {
    double fx, fy, cx, cy;
    Pose16d pose;
    mdvt_lookat look;
    double *out;
    mdvt_view *view;
    unsigned int *counter;  // zero on entry, left zero
};

__device__ __forceinline__ void finish_lookat(const double *partial, int n_blocks, const LookAtFinish &A) {
    const double fx = A.fx, fy = A.fy, cx = A.cx, cy = A.cy;
    const Pose16d &pose = A.pose;
    const mdvt_lookat &look = A.look;
    double *out = A.out;
    mdvt_view *view = A.view;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < n_blocks; b += kThreads) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += __ldcg(partial + b * 4 + k);
    }
    block_sum4(acc);
    if (threadIdx.x != 0) return;
    const double sz = acc[2], cnt = acc[3];
    double X = (acc[0] - cx * sz) / fx, Y = (acc[1] - cy * sz) / fy, Z = sz;
    if (pose.on) {
        const double *m = pose.m;
        const double x2 = m[0] * X + m[1] * Y + m[2] * Z + m[3] * cnt;
        const double y2 = m[4] * X + m[5] * Y + m[6] * Z + m[7] * cnt;
        const double z2 = m[8] * X + m[9] * Y + m[10] * Z + m[11] * cnt;
        X = x2; Y = y2; Z = z2;
    }
    out[0] = X; out[1] = Y; out[2] = Z; out[3] = cnt;
    double t[3] = {X / cnt, Y / cnt, Z / cnt};
#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (look.target_set[a]) t[a] = look.target[a];
    double f[3] = {t[0] - look.cam_pos[0], t[1] - look.cam_pos[1], t[2] - look.cam_pos[2]};
    const double nf = sqrt((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]);
    f[0] /= nf; f[1] /= nf; f[2] /= nf;
    double r[3] = {1.0 * f[2] - 0.0 * f[1], 0.0 * f[0] - 0.0 * f[2], 0.0 * f[1] - 1.0 * f[0]};  // cross(up = (0, 1, 0), f)
    const double nr = sqrt((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]);
    r[0] /= nr; r[1] /= nr; r[2] /= nr;
    const double u[3] = {f[1] * r[2] - f[2] * r[1], f[2] * r[0] - f[0] * r[2], f[0] * r[1] - f[1] * r[0]};
    double M[12] = {r[0], u[0] * look.y_scale, f[0], look.cam_pos[0],
                    r[1], u[1] * look.y_scale, f[1], look.cam_pos[1],
                    r[2], u[2] * look.y_scale, f[2], -look.cam_pos[2]};
    if (pose.on) {  // M (3x4) @ pose (4x4)
        double P[12];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 4; ++j)
                P[i * 4 + j] = ((M[i * 4] * pose.m[j] + M[i * 4 + 1] * pose.m[4 + j]) + M[i * 4 + 2] * pose.m[8 + j]) + M[i * 4 + 3] * pose.m[12 + j];
        for (int k = 0; k < 12; ++k) M[k] = P[k];
    }
    for (int k = 0; k < 12; ++k) view->M[k] = (float)M[k];
    view->fx = look.fx; view->fy = look.fy; view->cx = look.cx; view->cy = look.cy;
}

// Vertex sums of one frame.  The per-vertex formula is X = (xg - cx) * z / fx (depth_map_tools.py:1127-1128); its
// SUM over the frame is rearranged so that the per-pixel work is float64 multiply-adds and no division:
//     sum X = (sum(xg * z) - cx * sum(z)) / fx,   sum Y likewise,   sum Z = sum(z)
// (the pose, being affine, is applied to the mean on the host side of the ABI: mean(T p) = T mean(p)).  The result
// differs from a sum of individually rounded vertices by O(1e-16) relative -- the reference's own mean is a pairwise
// float64 sum, i.e. no bit-exact target exists; tests hold it to 1e-12 relative.
// partial[block] = {sum(xg*z), sum(yg*z), sum(z), count}.
// VEC4 (width % 4 == 0, word-aligned frame): a thread takes groups of 4 consecutive pixels of one row from three
// aligned words, grid-strided; (row, col) advance by a constant step with a carry, so there is no division in the
// loop, yg multiplies the group's depth sum once, and two groups are in flight per iteration (6 independent loads).
// v1 of this kernel spent 52 instructions per pixel (64-bit div/mod per pixel): 24.5 us per 4K frame at 12 % of HBM.
template <int DECODER, bool BIT16>
__device__ __forceinline__ void group_depths(const uint32_t w0, const uint32_t w1, const uint32_t w2, float dec_const, float depth_scale,
                                             double (&z)[4]) {
    const uint32_t r[4] = {w0 & 0xFF, w0 >> 24, (w1 >> 16) & 0xFF, (w2 >> 8) & 0xFF};
    const uint32_t gr[4] = {(w0 >> 8) & 0xFF, w1 & 0xFF, w1 >> 24, (w2 >> 16) & 0xFF};
    const uint32_t b[4] = {(w0 >> 16) & 0xFF, (w1 >> 8) & 0xFF, w2 & 0xFF, w2 >> 24};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        z[k] = (double)__fmul_rn(depth_of<DECODER>(code_of<DECODER, BIT16>(r[k], gr[k], b[k]), dec_const), depth_scale);
}

template <int DECODER, bool BIT16, bool VEC4, bool LOOKAT>
__global__ void __launch_bounds__(kThreads)
    centroid_partial_kernel(const void *__restrict__ src, int width, int64_t n, float dec_const, float depth_scale, float sx, float sy,
                            int stretched, double *partial, LookAtFinish fin) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (VEC4) {
        const int gpr = width / 4;                                   // groups per row
        const uint32_t n4 = (uint32_t)(n / 4), step = gridDim.x * kThreads;
        const int step_row = (int)(step / gpr), step_col = (int)(step - (uint32_t)step_row * gpr);
        uint32_t g = blockIdx.x * kThreads + threadIdx.x;
        int row = (int)(g / gpr), gc = (int)(g - (uint32_t)row * gpr);
        auto xg_of = [&](int col) { return (double)__fmul_rn(__int2float_rn(col), sx); };
        auto yg_of = [&](int r) { return stretched ? (double)__fmul_rn(__int2float_rn(r), sy) : (double)r; };
        auto accumulate = [&](const double (&z)[4], int r, int c4) {
            const int col = c4 * 4;
            double x0, x1, x2, x3;
            if (stretched) {
                x0 = xg_of(col); x1 = xg_of(col + 1); x2 = xg_of(col + 2); x3 = xg_of(col + 3);
            } else {  // exact small integers: one conversion per group
                x0 = (double)col; x1 = x0 + 1.0; x2 = x0 + 2.0; x3 = x0 + 3.0;
            }
            acc[0] = fma(x0, z[0], acc[0]);
            acc[0] = fma(x1, z[1], acc[0]);
            acc[0] = fma(x2, z[2], acc[0]);
            acc[0] = fma(x3, z[3], acc[0]);
            const double zs = (z[0] + z[1]) + (z[2] + z[3]);
            acc[1] = fma(yg_of(r), zs, acc[1]);
            acc[2] += zs;
        };
        auto advance = [&](int &r, int &c4) {
            c4 += step_col; r += step_row;
            if (c4 >= gpr) { c4 -= gpr; ++r; }
        };
        double cnt = 0.0;
        while (g < n4) {
            int row2 = row, gc2 = gc;
            advance(row2, gc2);
            const uint32_t g2 = g + step;
            const bool two = g2 < n4;
            uint32_t a0, a1, a2, b0 = 0, b1 = 0, b2 = 0;
            if (DECODER == MDVT_SOURCE_F32) {
                const float4 fa = __ldg(reinterpret_cast<const float4 *>(src) + g);
                float4 fb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (two) fb = __ldg(reinterpret_cast<const float4 *>(src) + g2);
                const double za[4] = {(double)__fmul_rn(fa.x, depth_scale), (double)__fmul_rn(fa.y, depth_scale),
                                      (double)__fmul_rn(fa.z, depth_scale), (double)__fmul_rn(fa.w, depth_scale)};
                accumulate(za, row, gc);
                if (two) {
                    const double zb[4] = {(double)__fmul_rn(fb.x, depth_scale), (double)__fmul_rn(fb.y, depth_scale),
                                          (double)__fmul_rn(fb.z, depth_scale), (double)__fmul_rn(fb.w, depth_scale)};
                    accumulate(zb, row2, gc2);
                }
            } else {
                const uint32_t *words = reinterpret_cast<const uint32_t *>(src);
                a0 = __ldg(words + 3 * (size_t)g); a1 = __ldg(words + 3 * (size_t)g + 1); a2 = __ldg(words + 3 * (size_t)g + 2);
                if (two) { b0 = __ldg(words + 3 * (size_t)g2); b1 = __ldg(words + 3 * (size_t)g2 + 1); b2 = __ldg(words + 3 * (size_t)g2 + 2); }
                double za[4];
                group_depths<DECODER, BIT16>(a0, a1, a2, dec_const, depth_scale, za);
                accumulate(za, row, gc);
                if (two) {
                    double zb[4];
                    group_depths<DECODER, BIT16>(b0, b1, b2, dec_const, depth_scale, zb);
                    accumulate(zb, row2, gc2);
                }
            }
            cnt += two ? 8.0 : 4.0;
            row = row2; gc = gc2;
            advance(row, gc);
            g = g2 + step;
        }
        acc[3] = cnt;
    } else {
        for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
            const int row = (int)(p / width), col = (int)(p - (int64_t)row * width);
            const double xg = stretched ? (double)__fmul_rn(__int2float_rn(col), sx) : (double)col;
            const double yg = stretched ? (double)__fmul_rn(__int2float_rn(row), sy) : (double)row;
            const double z = (double)__fmul_rn(source_depth<DECODER, BIT16>(src, p, dec_const), depth_scale);
            acc[0] = fma(xg, z, acc[0]);
            acc[1] = fma(yg, z, acc[1]);
            acc[2] += z;
            acc[3] += 1.0;
        }
    }
    block_sum4(acc);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) partial[blockIdx.x * 4 + k] = acc[k];
    }
    if (LOOKAT) {
        __shared__ int is_last;
        if (threadIdx.x == 0) {
            __threadfence();
            is_last = atomicAdd(fin.counter, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            finish_lookat(partial, (int)gridDim.x, fin);
            if (threadIdx.x == 0) *fin.counter = 0u;
        }
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
    const int64_t cap = sm_count() * 6 < kMaxBlocks - 1 ? sm_count() * 6 : kMaxBlocks - 1;  // the last scratch double is left free
    return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace mdvt

using namespace mdvt;

static int launch_centroid_partials(const void *depth_src, const mdvt_source *src, double *partial, int *grid_out, const LookAtFinish *fin,
                                    cudaStream_t st) {
    const int64_t n = (int64_t)src->width * src->height;
    const int stretched = !(src->grid_sx == 1.0f && src->grid_sy == 1.0f);
    const bool vec4 = (src->width % 4 == 0) && (n / 4 < 0x7FFFFFFFll) &&
                      (reinterpret_cast<uintptr_t>(depth_src) % (src->decoder == MDVT_SOURCE_F32 ? 16 : 4) == 0);
    const int grid = reduce_grid(vec4 ? n / 4 : n);
    const LookAtFinish none{};
#define ARGS depth_src, src->width, n, src->dec_const, src->depth_scale, src->grid_sx, src->grid_sy, stretched, partial
#define CALL(D, B)                                                                                                            \
    do {                                                                                                                      \
        if (vec4 && fin) centroid_partial_kernel<D, B, true, true><<<grid, kThreads, 0, st>>>(ARGS, *fin);                    \
        else if (vec4) centroid_partial_kernel<D, B, true, false><<<grid, kThreads, 0, st>>>(ARGS, none);                     \
        else if (fin) centroid_partial_kernel<D, B, false, true><<<grid, kThreads, 0, st>>>(ARGS, *fin);                      \
        else centroid_partial_kernel<D, B, false, false><<<grid, kThreads, 0, st>>>(ARGS, none);                              \
    } while (0)
    MDVT_DISPATCH_SOURCE(src->decoder, src->bit16, CALL);
#undef CALL
#undef ARGS
    MDVT_CUDA_TRY(cudaGetLastError());
    *grid_out = grid;
    return MDVT_OK;
}

extern "C" int mdvt_centroid(const void *depth_src, const mdvt_source *src, const double *K_host, const double *pose_host,
                             double *out_sums, void *stream) {
    if (int rc = check_source(src)) return rc;
    MDVT_REQUIRE(depth_src && K_host && out_sums, "NULL buffer");
    Pose12d pose{};
    if (pose_host) {
        for (int k = 0; k < 12; ++k) pose.m[k] = pose_host[k];
        pose.on = 1;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double *partial = out_sums + 4;
    int grid = 0;
    if (int rc = launch_centroid_partials(depth_src, src, partial, &grid, nullptr, st)) return rc;
    finish_centroid_kernel<<<1, kThreads, 0, st>>>(partial, grid, K_host[0], K_host[1], K_host[2], K_host[3], pose, out_sums);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

namespace mdvt {
// Centroid of one frame -> look-at camera of the same frame, all on the device (used by mdvt_novel_view_frames).
int launch_centroid_lookat(const void *depth_src, const mdvt_source *src, const double *K_host, const double *pose16_host,
                           const mdvt_lookat *look, double *out_sums, mdvt_view *view_dev, cudaStream_t st) {
    LookAtFinish fin{};
    fin.fx = K_host[0]; fin.fy = K_host[1]; fin.cx = K_host[2]; fin.cy = K_host[3];
    if (pose16_host) {
        for (int k = 0; k < 16; ++k) fin.pose.m[k] = pose16_host[k];
        fin.pose.on = 1;
    }
    fin.look = *look;
    fin.out = out_sums;
    fin.view = view_dev;
    fin.counter = reinterpret_cast<unsigned int *>(out_sums + 4 + MDVT_REDUCE_SCRATCH_DOUBLES - 1);  // last scratch double (never a partial)
    int grid = 0;
    return launch_centroid_partials(depth_src, src, out_sums + 4, &grid, &fin, st);
}
}  // namespace mdvt

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
