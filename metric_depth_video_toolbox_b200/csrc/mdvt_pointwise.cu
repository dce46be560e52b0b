// Wire-format codec and decode+unproject kernels (HBM-bound, one pass, no reuse).
#include "mdvt_common.cuh"

namespace mdvt {

constexpr int kThreads = 256;
constexpr int kPxPerThread = 4;  // 4 px = 12 input bytes = three aligned 32-bit words

// Load the 12 bytes of 4 consecutive pixels starting at pixel index p4*4 as three words and split
// them into r/g/b of each pixel.  Base pointers from the callers are >= 16-byte aligned and a
// 4-pixel group starts at byte 12*p4, so the word loads are aligned.
__device__ __forceinline__ void load_px4(const uint8_t *__restrict__ rgb, int64_t p4, uint32_t (&r)[4], uint32_t (&g)[4],
                                         uint32_t (&b)[4]) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(rgb) + p4 * 3;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    r[0] = w0 & 0xFF;         g[0] = (w0 >> 8) & 0xFF;  b[0] = (w0 >> 16) & 0xFF;
    r[1] = w0 >> 24;          g[1] = w1 & 0xFF;         b[1] = (w1 >> 8) & 0xFF;
    r[2] = (w1 >> 16) & 0xFF; g[2] = w1 >> 24;          b[2] = w2 & 0xFF;
    r[3] = (w2 >> 8) & 0xFF;  g[3] = (w2 >> 16) & 0xFF; b[3] = w2 >> 24;
}

__device__ __forceinline__ void load_px1(const uint8_t *__restrict__ rgb, int64_t p, uint32_t &r, uint32_t &g, uint32_t &b) {
    r = rgb[p * 3];
    g = rgb[p * 3 + 1];
    b = rgb[p * 3 + 2];
}

// ---------------------------------------------------------------------------------------------
template <int DECODER, bool BIT16>
__global__ void __launch_bounds__(kThreads) decode_kernel(const uint8_t *__restrict__ rgb, int64_t n, float dec_const,
                                                          uint32_t *__restrict__ out_codes, float *__restrict__ out_depth) {
    const int64_t n4 = n / kPxPerThread;
    for (int64_t p4 = blockIdx.x * (int64_t)kThreads + threadIdx.x; p4 < n4; p4 += (int64_t)gridDim.x * kThreads) {
        uint32_t r[4], g[4], b[4];
        load_px4(rgb, p4, r, g, b);
        uint4 c;
        c.x = code_of<DECODER, BIT16>(r[0], g[0], b[0]);
        c.y = code_of<DECODER, BIT16>(r[1], g[1], b[1]);
        c.z = code_of<DECODER, BIT16>(r[2], g[2], b[2]);
        c.w = code_of<DECODER, BIT16>(r[3], g[3], b[3]);
        if (out_codes) reinterpret_cast<uint4 *>(out_codes)[p4] = c;
        if (out_depth) {
            float4 d;
            d.x = depth_of<DECODER>(c.x, dec_const);
            d.y = depth_of<DECODER>(c.y, dec_const);
            d.z = depth_of<DECODER>(c.z, dec_const);
            d.w = depth_of<DECODER>(c.w, dec_const);
            reinterpret_cast<float4 *>(out_depth)[p4] = d;
        }
    }
    // ragged tail (n % 4 pixels), one thread each
    const int64_t p = n4 * kPxPerThread + blockIdx.x * (int64_t)kThreads + threadIdx.x;
    if (p < n) {
        uint32_t r, g, b;
        load_px1(rgb, p, r, g, b);
        const uint32_t c = code_of<DECODER, BIT16>(r, g, b);
        if (out_codes) out_codes[p] = c;
        if (out_depth) out_depth[p] = depth_of<DECODER>(c, dec_const);
    }
}

// ---------------------------------------------------------------------------------------------
// depth_frames_helper.py:5-11: np.clip in float32, float64 multiply, truncating cast; :48-61 bytes.
// T = the array's own dtype: np.clip works in it (float32 depth: max_depth rounded to float32; float64 depth: exact).
template <typename T>
__global__ void __launch_bounds__(kThreads) encode_kernel(const T *__restrict__ depth, int64_t n, T max_depth_t,
                                                          double multiplier, int bit16, int bgr_order,
                                                          uint32_t *__restrict__ out_codes, uint8_t *__restrict__ out_pix) {
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        T d = depth[p];
        // np.clip == minimum(maximum(d, 0), max): NaN propagates; a NaN code is written as 0
        d = d < (T)0 ? (T)0 : d;
        d = d > max_depth_t ? max_depth_t : d;
        const double scaled = __dmul_rn(multiplier, (double)d);
        const uint32_t code = (d == d) ? __double2uint_rz(scaled) : 0u;
        if (out_codes) out_codes[p] = code;
        if (out_pix) {
            uint8_t c0, c1, c2;  // R, G, B
            if (bit16) {
                c0 = c1 = (uint8_t)(code >> 24);
                c2 = (uint8_t)(code >> 16);
            } else {
                c0 = (uint8_t)(code >> 16);
                c1 = (uint8_t)(code >> 8);
                c2 = (uint8_t)code;
            }
            uint8_t *o = out_pix + p * 3;
            if (bgr_order) {
                o[0] = c2; o[1] = c1; o[2] = c0;
            } else {
                o[0] = c0; o[1] = c1; o[2] = c2;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct Pose12f {
    float m[12];
    int on;
};
struct Pose12d {
    double m[12];
    int on;
};

template <int DECODER, bool BIT16>
__global__ void __launch_bounds__(kThreads)
    unproject_f32_kernel(const void *__restrict__ rgb, int width, int64_t n, float dec_const, float depth_scale, SourceCam cam,
                         Pose12f pose, float *__restrict__ out_xyz) {
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        const float z = __fmul_rn(source_depth<DECODER, BIT16>(rgb, p, dec_const), depth_scale);
        const int row = (int)(p / width), col = (int)(p - (int64_t)row * width);
        float X, Y, Z = z;
        unproject_px(cam, col, row, z, X, Y);
        if (pose.on) {
            const float x2 = affine_row(pose.m, X, Y, Z), y2 = affine_row(pose.m + 4, X, Y, Z), z2 = affine_row(pose.m + 8, X, Y, Z);
            X = x2; Y = y2; Z = z2;
        }
        out_xyz[p * 3] = X;
        out_xyz[p * 3 + 1] = Y;
        out_xyz[p * 3 + 2] = Z;
    }
}

// float64 twin with NumPy's evaluation order: the grid stretch is a float32 multiply (:1120-1123),
// everything after the subtraction is float64 (NumPy >= 2 promotion).
template <int DECODER, bool BIT16>
__global__ void __launch_bounds__(kThreads)
    unproject_f64_kernel(const void *__restrict__ rgb, int width, int64_t n, float dec_const, float depth_scale, float sx,
                         float sy, int stretched, double fx, double fy, double cx, double cy, Pose12d pose,
                         double *__restrict__ out_xyz) {
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        const float zf = __fmul_rn(source_depth<DECODER, BIT16>(rgb, p, dec_const), depth_scale);
        const int row = (int)(p / width), col = (int)(p - (int64_t)row * width);
        double xg = (double)col, yg = (double)row;
        if (stretched) {
            xg = (double)__fmul_rn(__int2float_rn(col), sx);
            yg = (double)__fmul_rn(__int2float_rn(row), sy);
        }
        const double z = (double)zf;
        double X = __ddiv_rn(__dmul_rn(__dsub_rn(xg, cx), z), fx);
        double Y = __ddiv_rn(__dmul_rn(__dsub_rn(yg, cy), z), fy);
        double Z = z;
        if (pose.on) {
            const double *m = pose.m;
            const double x2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[0], X), __dmul_rn(m[1], Y)), __dmul_rn(m[2], Z)), m[3]);
            const double y2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[4], X), __dmul_rn(m[5], Y)), __dmul_rn(m[6], Z)), m[7]);
            const double z2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[8], X), __dmul_rn(m[9], Y)), __dmul_rn(m[10], Z)), m[11]);
            X = x2; Y = y2; Z = z2;
        }
        out_xyz[p * 3] = X;
        out_xyz[p * 3 + 1] = Y;
        out_xyz[p * 3 + 2] = Z;
    }
}

// depth_map_tools.transform_points (:977-1004): [x y z 1] @ T.T, w dropped without a divide.  NumPy's matmul
// accumulates left to right over the 4 terms of each output; reproduced with separately rounded operations.
__global__ void __launch_bounds__(kThreads) transform_points_f64_kernel(const double *__restrict__ xyz, int64_t n, Pose12d pose,
                                                                        double *__restrict__ out) {
    const double *m = pose.m;
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        const double X = xyz[p * 3], Y = xyz[p * 3 + 1], Z = xyz[p * 3 + 2];
        out[p * 3] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(X, m[0]), __dmul_rn(Y, m[1])), __dmul_rn(Z, m[2])), m[3]);
        out[p * 3 + 1] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(X, m[4]), __dmul_rn(Y, m[5])), __dmul_rn(Z, m[6])), m[7]);
        out[p * 3 + 2] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(X, m[8]), __dmul_rn(Y, m[9])), __dmul_rn(Z, m[10])), m[11]);
    }
}

// depth_map_tools.project_3d_points_to_2d (:1057-1060) without distortion: u = fx X/Z + cx, v = fy Y/Z + cy in
// float64 (cv2.projectPoints with zero rvec/tvec/dist; its Z == 0 -> 1/z := 1 rule is kept).
__global__ void __launch_bounds__(kThreads) project_points_f64_kernel(const double *__restrict__ xyz, int64_t n, double fx, double fy,
                                                                      double cx, double cy, double *__restrict__ out_uv) {
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        const double X = xyz[p * 3], Y = xyz[p * 3 + 1], Z = xyz[p * 3 + 2];
        const double iz = Z != 0.0 ? __ddiv_rn(1.0, Z) : 1.0;
        out_uv[p * 2] = __dadd_rn(__dmul_rn(fx, __dmul_rn(X, iz)), cx);
        out_uv[p * 2 + 1] = __dadd_rn(__dmul_rn(fy, __dmul_rn(Y, iz)), cy);
    }
}

static int grid_for(int64_t work_items) {
    const int64_t blocks = (work_items + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
    return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

// decode_uint32_as_depth (depth_frames_helper.py:13-24) and the inline divide variants on codes the caller holds.
template <int DECODER>
__global__ void __launch_bounds__(kThreads) codes_to_depth_kernel(const uint32_t *__restrict__ codes, int64_t n, float dec_const,
                                                                  float *__restrict__ out_depth) {
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads)
        out_depth[p] = depth_of<DECODER>(__ldg(codes + p), dec_const);
}

// encode_data_as_BGR (depth_frames_helper.py:48-61): bytes of a u32 plane -> u8x3.
__global__ void __launch_bounds__(kThreads) codes_to_pixels_kernel(const uint32_t *__restrict__ codes, int64_t n, int bit16, int bgr_order,
                                                                   uint8_t *__restrict__ out_pix) {
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        const uint32_t code = __ldg(codes + p);
        uint8_t c0, c1, c2;  // R, G, B
        if (bit16) {
            c0 = c1 = (uint8_t)(code >> 24);
            c2 = (uint8_t)(code >> 16);
        } else {
            c0 = (uint8_t)(code >> 16);
            c1 = (uint8_t)(code >> 8);
            c2 = (uint8_t)code;
        }
        uint8_t *o = out_pix + p * 3;
        if (bgr_order) {
            o[0] = c2; o[1] = c1; o[2] = c0;
        } else {
            o[0] = c0; o[1] = c1; o[2] = c2;
        }
    }
}

// depth_map_tools.calculate_normals (depth_map_tools.py:20-60) for a float32 depth plane, float32 like NumPy computes it
// there (float32 arrays against Python scalars stay float32): P = ((u - cx)/fx * z, (cy - v)/fy * z, z), forward
// differences to the right / lower neighbour (the last column / row repeats itself: a zero difference), their cross
// product (multiply, multiply, subtract per component), divided by sqrt((n0^2 + n1^2) + n2^2) + 1e-8, then y and z
// negated ("DirectX conversion").
__global__ void __launch_bounds__(kThreads)
    calculate_normals_kernel(const float *__restrict__ depth, int width, int height, float fx, float fy, float cx, float cy,
                             float *__restrict__ out) {
    const int64_t n = (int64_t)width * height;
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        const int row = (int)(p / width), col = (int)(p - (int64_t)row * width);
        auto point = [&](int c, int r, float &X, float &Y, float &Z) {
            Z = __ldg(depth + (int64_t)r * width + c);
            X = __fmul_rn(__fdiv_rn(__fsub_rn((float)c, cx), fx), Z);
            Y = __fmul_rn(__fdiv_rn(__fsub_rn(cy, (float)r), fy), Z);
        };
        float x0, y0, z0, xr, yr, zr, xd, yd, zd;
        point(col, row, x0, y0, z0);
        point(min(col + 1, width - 1), row, xr, yr, zr);
        point(col, min(row + 1, height - 1), xd, yd, zd);
        const float ax = __fsub_rn(xr, x0), ay = __fsub_rn(yr, y0), az = __fsub_rn(zr, z0);
        const float bx = __fsub_rn(xd, x0), by = __fsub_rn(yd, y0), bz = __fsub_rn(zd, z0);
        const float n0 = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
        const float n1 = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
        const float n2 = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
        const float len = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(n0, n0), __fmul_rn(n1, n1)), __fmul_rn(n2, n2))), 1e-8f);
        float *o = out + p * 3;
        o[0] = __fdiv_rn(n0, len);
        o[1] = -__fdiv_rn(n1, len);
        o[2] = -__fdiv_rn(n2, len);
    }
}

}  // namespace mdvt

using namespace mdvt;

extern "C" int mdvt_decode_depth(const uint8_t *rgb, int64_t n_pixels, int decoder, int bit16, float dec_const,
                                 uint32_t *out_codes, float *out_depth, void *stream) {
    MDVT_REQUIRE(n_pixels >= 0, "negative pixel count");
    if (int rc = check_decoder(decoder, bit16, false)) return rc;
    if (n_pixels == 0 || (!out_codes && !out_depth)) return MDVT_OK;
    MDVT_REQUIRE(rgb != nullptr, "rgb is NULL");
    MDVT_REQUIRE((reinterpret_cast<uintptr_t>(rgb) & 3) == 0, "rgb must be 4-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for((n_pixels + kPxPerThread - 1) / kPxPerThread);
#define CALL(D, B) decode_kernel<D, B><<<grid, kThreads, 0, st>>>(rgb, n_pixels, dec_const, out_codes, out_depth)
    MDVT_DISPATCH_SOURCE(decoder, bit16, CALL);
#undef CALL
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_encode_depth(const float *depth, int64_t n_pixels, double max_depth, int bit16, int bgr_order,
                                 uint32_t *out_codes, uint8_t *out_pix, void *stream) {
    MDVT_REQUIRE(n_pixels >= 0, "negative pixel count");
    MDVT_REQUIRE(max_depth > 0, "max_depth must be positive");
    if (n_pixels == 0 || (!out_codes && !out_pix)) return MDVT_OK;
    MDVT_REQUIRE(depth != nullptr, "depth is NULL");
    const double multiplier = 4228250625.0 / max_depth;  // 255**4 / float(max_depth)
    encode_kernel<float><<<grid_for(n_pixels), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        depth, n_pixels, (float)max_depth, multiplier, bit16, bgr_order, out_codes, out_pix);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_encode_depth_f64(const double *depth, int64_t n_pixels, double max_depth, int bit16, int bgr_order,
                                     uint32_t *out_codes, uint8_t *out_pix, void *stream) {
    MDVT_REQUIRE(n_pixels >= 0, "negative pixel count");
    MDVT_REQUIRE(max_depth > 0, "max_depth must be positive");
    if (n_pixels == 0 || (!out_codes && !out_pix)) return MDVT_OK;
    MDVT_REQUIRE(depth != nullptr, "depth is NULL");
    encode_kernel<double><<<grid_for(n_pixels), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        depth, n_pixels, max_depth, 4228250625.0 / max_depth, bit16, bgr_order, out_codes, out_pix);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_unproject_f32(const void *depth_rgb, const mdvt_source *src, const float *pose_host, float *out_xyz,
                                  void *stream) {
    if (int rc = check_source(src)) return rc;
    MDVT_REQUIRE(depth_rgb && out_xyz, "NULL buffer");
    const int64_t n = (int64_t)src->width * src->height;
    SourceCam cam{src->fx, src->fy, src->cx, src->cy, src->grid_sx, src->grid_sy};
    Pose12f pose{};
    if (pose_host) {
        for (int k = 0; k < 12; ++k) pose.m[k] = pose_host[k];
        pose.on = 1;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(D, B) \
    unproject_f32_kernel<D, B><<<grid_for(n), kThreads, 0, st>>>(depth_rgb, src->width, n, src->dec_const, src->depth_scale, cam, pose, out_xyz)
    MDVT_DISPATCH_SOURCE(src->decoder, src->bit16, CALL);
#undef CALL
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_unproject_f64(const void *depth_rgb, const mdvt_source *src, const double *K_host,
                                  const double *pose_host, double *out_xyz, void *stream) {
    if (int rc = check_source(src)) return rc;
    MDVT_REQUIRE(depth_rgb && out_xyz && K_host, "NULL buffer");
    const int64_t n = (int64_t)src->width * src->height;
    Pose12d pose{};
    if (pose_host) {
        for (int k = 0; k < 12; ++k) pose.m[k] = pose_host[k];
        pose.on = 1;
    }
    const int stretched = !(src->grid_sx == 1.0f && src->grid_sy == 1.0f);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(D, B)                                                                                                          \
    unproject_f64_kernel<D, B><<<grid_for(n), kThreads, 0, st>>>(depth_rgb, src->width, n, src->dec_const, src->depth_scale, \
                                                                 src->grid_sx, src->grid_sy, stretched, K_host[0], K_host[1], \
                                                                 K_host[2], K_host[3], pose, out_xyz)
    MDVT_DISPATCH_SOURCE(src->decoder, src->bit16, CALL);
#undef CALL
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_codes_to_depth(const uint32_t *codes, int64_t n_pixels, int decoder, float dec_const, float *out_depth, void *stream) {
    MDVT_REQUIRE(n_pixels >= 0, "negative pixel count");
    if (int rc = check_decoder(decoder, 1, false)) return rc;
    if (n_pixels == 0) return MDVT_OK;
    MDVT_REQUIRE(codes && out_depth, "NULL buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (decoder == MDVT_DECODE_D1)
        codes_to_depth_kernel<MDVT_DECODE_D1><<<grid_for(n_pixels), kThreads, 0, st>>>(codes, n_pixels, dec_const, out_depth);
    else
        codes_to_depth_kernel<MDVT_DECODE_D3><<<grid_for(n_pixels), kThreads, 0, st>>>(codes, n_pixels, dec_const, out_depth);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_codes_to_pixels(const uint32_t *codes, int64_t n_pixels, int bit16, int bgr_order, uint8_t *out_pix, void *stream) {
    MDVT_REQUIRE(n_pixels >= 0, "negative pixel count");
    if (n_pixels == 0) return MDVT_OK;
    MDVT_REQUIRE(codes && out_pix, "NULL buffer");
    codes_to_pixels_kernel<<<grid_for(n_pixels), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(codes, n_pixels, bit16, bgr_order, out_pix);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_transform_points_f64(const double *xyz, int64_t n_points, const double *pose_host, double *out_xyz, void *stream) {
    MDVT_REQUIRE(n_points >= 0, "negative point count");
    MDVT_REQUIRE(pose_host != nullptr, "pose is NULL");
    if (n_points == 0) return MDVT_OK;
    MDVT_REQUIRE(xyz && out_xyz, "NULL buffer");
    Pose12d pose{};
    for (int k = 0; k < 12; ++k) pose.m[k] = pose_host[k];
    pose.on = 1;
    transform_points_f64_kernel<<<grid_for(n_points), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(xyz, n_points, pose, out_xyz);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_project_points_f64(const double *xyz, int64_t n_points, const double *K_host, double *out_uv, void *stream) {
    MDVT_REQUIRE(n_points >= 0, "negative point count");
    MDVT_REQUIRE(K_host != nullptr, "K is NULL");
    if (n_points == 0) return MDVT_OK;
    MDVT_REQUIRE(xyz && out_uv, "NULL buffer");
    project_points_f64_kernel<<<grid_for(n_points), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(xyz, n_points, K_host[0], K_host[1],
                                                                                                      K_host[2], K_host[3], out_uv);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_calculate_normals(const float *depth, int width, int height, const double *K_host, float *out_normals, void *stream) {
    MDVT_REQUIRE(width > 0 && height > 0, "bad frame size %dx%d", width, height);
    MDVT_REQUIRE(depth && K_host && out_normals, "NULL buffer");
    // the reference reads K through float(): Python doubles against float32 arrays -> NumPy computes in float32 with the
    // scalar rounded to float32
    calculate_normals_kernel<<<grid_for((int64_t)width * height), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        depth, width, height, (float)K_host[0], (float)K_host[1], (float)K_host[2], (float)K_host[3], out_normals);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}
