// Normals-coded infill mask, per-pixel parts:
//   E1  the mesh builder's edge test on the depth grid (depth_map_tools.py:1243-1376): which vertices belong to a
//       triangle that is seen at more than 89 degrees, and the unit normal the reference attaches to each of them
//   E2  those "edge points" moved into the eye camera and z-buffered (stereo_rerender.py:589-606,727-735,745-755)
//   E3  per target pixel: hole / border / edge-normal colour of the mask image and the edge colour painted into the
//       eye image (:787-803,813-814)
// The TELEA inpainting and masked blur that follow (:805-808) are OpenCV calls on the host, as in the reference.
// Everything geometric here is float64 with the reference's evaluation order (the reference computes it in float64
// and thresholds / truncates the results, so float32 would flip decisions).
#include <cmath>

#include "mdvt_common.cuh"

namespace mdvt {

constexpr int kThreads = 256;

struct Cam64 {
    double fx, fy, cx, cy;
    float sx, sy;   // of_by_one grid stretch (float32 multiply, then promoted: depth_map_tools.py:1118-1123)
    int stretched;
};

struct Pose12d {
    double m[12];
};

struct V3 {
    double x, y, z;
};
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ double dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ V3 apply(const Pose12d &p, V3 v) {
    const double *m = p.m;  // points @ R.T + t, summed left to right
    return {((v.x * m[0] + v.y * m[1]) + v.z * m[2]) + m[3], ((v.x * m[4] + v.y * m[5]) + v.z * m[6]) + m[7],
            ((v.x * m[8] + v.y * m[9]) + v.z * m[10]) + m[11]};
}

// Mesh vertex (row, col): create_point_cloud_from_depth on the (optionally stretched) grid, float64 -- or, DECODER ==
// kSourceXYZ64, read from the (H*W, 3) float64 points the caller of create_mesh_from_point_cloud holds.
constexpr int kSourceXYZ64 = 100;
template <int DECODER, bool BIT16>
__device__ __forceinline__ V3 vertex_at(const void *__restrict__ src, int width, int row, int col, float dec_const, float depth_scale,
                                        const Cam64 &c) {
    if (DECODER == kSourceXYZ64) {
        const double *p = reinterpret_cast<const double *>(src) + ((int64_t)row * width + col) * 3;
        return {p[0], p[1], p[2]};
    }
    const float zf = __fmul_rn(source_depth<DECODER, BIT16>(src, (int64_t)row * width + col, dec_const), depth_scale);
    const double xg = c.stretched ? (double)__fmul_rn(__int2float_rn(col), c.sx) : (double)col;
    const double yg = c.stretched ? (double)__fmul_rn(__int2float_rn(row), c.sy) : (double)row;
    const double z = (double)zf;
    return {(xg - c.cx) * z / c.fx, (yg - c.cy) * z / c.fy, z};
}

// Raw normal of a triangle and its 89-degree verdict (depth_map_tools.py:1283-1294).
__device__ __forceinline__ bool triangle_bad(V3 v1, V3 v2, V3 v3, double cos_limit, V3 &normal) {
    normal = cross3(v2 - v1, v3 - v1);
    const V3 s = (v1 + v2) + v3;
    const V3 view = {-s.x / 3.0, -s.y / 3.0, -s.z / 3.0};
    const double cosine = dot3(normal, view) / (sqrt(dot3(normal, normal)) * sqrt(dot3(view, view)) + 1e-15);
    return cosine < cos_limit;
}

// E1a: one thread per grid cell -> bit 0: triangle A bad, bit 1: triangle B bad.
template <int DECODER, bool BIT16>
__global__ void __launch_bounds__(kThreads)
    edge_cells_kernel(const void *__restrict__ src, int width, int height, float dec_const, float depth_scale, Cam64 cam, double cos_limit,
                      uint8_t *__restrict__ cell_flags) {
    const int cw = width - 1, ch = height - 1;
    const int64_t n = (int64_t)cw * ch;
    for (int64_t c = blockIdx.x * (int64_t)kThreads + threadIdx.x; c < n; c += (int64_t)gridDim.x * kThreads) {
        const int i = (int)(c / cw), j = (int)(c - (int64_t)i * cw);
        const V3 p00 = vertex_at<DECODER, BIT16>(src, width, i, j, dec_const, depth_scale, cam);
        const V3 p10 = vertex_at<DECODER, BIT16>(src, width, i + 1, j, dec_const, depth_scale, cam);
        const V3 p11 = vertex_at<DECODER, BIT16>(src, width, i + 1, j + 1, dec_const, depth_scale, cam);
        const V3 p01 = vertex_at<DECODER, BIT16>(src, width, i, j + 1, dec_const, depth_scale, cam);
        V3 n_unused;
        const bool bad_a = triangle_bad(p00, p10, p11, cos_limit, n_unused);
        const bool bad_b = triangle_bad(p00, p11, p01, cos_limit, n_unused);
        cell_flags[c] = (uint8_t)((bad_a ? 1 : 0) | (bad_b ? 2 : 0));
    }
}

// E1b: one thread per vertex.  A vertex is "unused" when any triangle listing it is bad (:1329-1335); its normal is the
// unit normal of the last triangle listing it in the order "all A row-major, then all B" (:1351-1361): B of its own
// cell; on the right column B of cell (i, W-2); on the bottom row B of cell (H-2, j-1); bottom-left corner: A of
// cell (H-2, 0).  A zero-area triangle gives (1, 1, 1) (:1339-1345).
template <int DECODER, bool BIT16>
__global__ void __launch_bounds__(kThreads)
    edge_vertices_kernel(const void *__restrict__ src, int width, int height, float dec_const, float depth_scale, Cam64 cam,
                         const uint8_t *__restrict__ cell_flags, uint8_t *__restrict__ out_flags, double *__restrict__ out_normals) {
    const int cw = width - 1, ch = height - 1;
    const int64_t n = (int64_t)width * height;
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        const int i = (int)(p / width), j = (int)(p - (int64_t)i * width);
        auto flag = [&](int ci, int cj) -> uint32_t {
            return (ci >= 0 && cj >= 0 && ci < ch && cj < cw) ? cell_flags[(int64_t)ci * cw + cj] : 0u;
        };
        // A lists (i,j) in cells (i,j), (i-1,j), (i-1,j-1); B in cells (i,j), (i-1,j-1), (i,j-1)
        const uint32_t own = flag(i, j), up = flag(i - 1, j), diag = flag(i - 1, j - 1), left = flag(i, j - 1);
        const bool unused = (own & 3u) || (up & 1u) || (diag & 3u) || (left & 2u);
        out_flags[p] = unused ? 1 : 0;
        if (!unused || !out_normals) continue;
        int ci, cj;
        bool use_a = false;
        if (i < ch && j < cw) { ci = i; cj = j; }
        else if (i < ch) { ci = i; cj = cw - 1; }
        else if (j >= 1) { ci = ch - 1; cj = j - 1; }
        else { ci = ch - 1; cj = 0; use_a = true; }
        const V3 p00 = vertex_at<DECODER, BIT16>(src, width, ci, cj, dec_const, depth_scale, cam);
        const V3 p11 = vertex_at<DECODER, BIT16>(src, width, ci + 1, cj + 1, dec_const, depth_scale, cam);
        const V3 third = use_a ? vertex_at<DECODER, BIT16>(src, width, ci + 1, cj, dec_const, depth_scale, cam)
                               : vertex_at<DECODER, BIT16>(src, width, ci, cj + 1, dec_const, depth_scale, cam);
        const V3 nrm = use_a ? cross3(third - p00, p11 - p00) : cross3(p11 - p00, third - p00);
        const double len = sqrt((nrm.x * nrm.x + nrm.y * nrm.y) + nrm.z * nrm.z);
        double *o = out_normals + p * 3;
        if (len > 0.0) { o[0] = nrm.x / len; o[1] = nrm.y / len; o[2] = nrm.z / len; }
        else { o[0] = 1.0; o[1] = 1.0; o[2] = 1.0; }
    }
}

// Frame-space edge point of vertex p and the end point of its normal (stereo_rerender.py:594-600): the end point is
// built on the stretched vertex, the point itself is then squeezed by (W-1)/W, (H-1)/H.
template <int DECODER, bool BIT16>
__device__ __forceinline__ void edge_point(const void *__restrict__ src, int width, int height, int64_t p, float dec_const,
                                           float depth_scale, const Cam64 &cam, const double *__restrict__ normals, double squeeze_x,
                                           double squeeze_y, V3 &point, V3 &end) {
    const int i = (int)(p / width), j = (int)(p - (int64_t)i * width);
    const V3 v = vertex_at<DECODER, BIT16>(src, width, i, j, dec_const, depth_scale, cam);
    if (normals) end = {normals[p * 3] + v.x, normals[p * 3 + 1] + v.y, normals[p * 3 + 2] + v.z};
    point = {v.x * squeeze_x, v.y * squeeze_y, v.z};
}

// E2: edge points -> eye camera -> rounded pixel -> 64-bit atomicMin (nearest wins, ties -> lowest vertex index).
template <int DECODER, bool BIT16>
__global__ void __launch_bounds__(kThreads)
    edge_splat_kernel(const void *__restrict__ src, int width, int height, float dec_const, float depth_scale, Cam64 cam,
                      const uint8_t *__restrict__ flags, Pose12d pose, double rfx, double rfy, double rcx, double rcy, double squeeze_x,
                      double squeeze_y, int out_w, int out_h, unsigned long long *__restrict__ zbuf) {
    const int64_t n = (int64_t)width * height;
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        if (!flags[p]) continue;
        V3 pt, end;
        edge_point<DECODER, BIT16>(src, width, height, p, dec_const, depth_scale, cam, nullptr, squeeze_x, squeeze_y, pt, end);
        const V3 q = apply(pose, pt);
        if (!(q.z > 0.0)) continue;  // behind the camera: cv2.projectPoints would mirror it; not a case the scripts produce
        const double u = rint(rfx * q.x / q.z + rcx), v = rint(rfy * q.y / q.z + rcy);  // np.round: half to even
        if (u >= 0.0 && u < (double)out_w && v >= 0.0 && v < (double)out_h) {
            const unsigned long long key = ((unsigned long long)__float_as_uint((float)q.z) << 32) | (uint32_t)p;
            atomicMin(zbuf + (int64_t)v * out_w + (int64_t)u, key);
        }
    }
}

// E3: one thread per target pixel of one eye.
template <int DECODER, bool BIT16>
__global__ void __launch_bounds__(kThreads)
    edge_resolve_kernel(unsigned long long *__restrict__ zbuf, const void *__restrict__ src, int width, int height, float dec_const,
                        float depth_scale, Cam64 cam, const double *__restrict__ normals, Pose12d pose, double squeeze_x, double squeeze_y,
                        const uint8_t *__restrict__ colour, const uint8_t *__restrict__ hole_mask, int64_t hole_pitch, int out_w, int out_h,
                        uint32_t bg_rgb, int code_normals, uint8_t *__restrict__ image, int64_t image_pitch, uint8_t *__restrict__ mask_img,
                        int64_t mask_pitch) {
    const int64_t n = (int64_t)out_w * out_h;
    for (int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x; t < n; t += (int64_t)gridDim.x * kThreads) {
        const int r = (int)(t / out_w), c = (int)(t - (int64_t)r * out_w);
        const unsigned long long key = zbuf[t];
        zbuf[t] = MDVT_ZBUF_EMPTY;
        const bool hole = hole_mask[r * hole_pitch + c] != 0;
        uint8_t m0 = 0, m1 = 0, m2 = 0;
        if (hole) {
            m0 = (uint8_t)bg_rgb; m1 = (uint8_t)(bg_rgb >> 8); m2 = (uint8_t)(bg_rgb >> 16);
            if (code_normals) {
                // border holes get a fixed inward normal; later rules only touch pixels that are still background
                // (stereo_rerender.py:796-799): column 0, column W-1, row 0, row H-1, in that order
                if (c == 0) { m0 = 255; m1 = 127; m2 = 127; }
                else if (c == out_w - 1) { m0 = 0; m1 = 127; m2 = 127; }
                else if (r == 0) { m0 = 127; m1 = 127; m2 = 0; }
                else if (r == out_h - 1) { m0 = 127; m1 = 127; m2 = 255; }
            }
            if (key != MDVT_ZBUF_EMPTY) {
                const int64_t p = (int64_t)(uint32_t)key;
                if (code_normals) {
                    V3 pt, end;
                    edge_point<DECODER, BIT16>(src, width, height, p, dec_const, depth_scale, cam, normals, squeeze_x, squeeze_y, pt, end);
                    const V3 d = apply(pose, end) - apply(pose, pt);
                    const double len = sqrt((d.x * d.x + d.y * d.y) + d.z * d.z);
                    m0 = (uint8_t)(int)(((d.x / len + 1.0) / 2.0) * 255.0);  // (normal + 1) / 2, later * 255 truncated (:802,817)
                    m1 = (uint8_t)(int)(((d.y / len + 1.0) / 2.0) * 255.0);
                    m2 = (uint8_t)(int)(((d.z / len + 1.0) / 2.0) * 255.0);
                }
                if (image) {  // the edge vertex's own colour fills the hole pixel (:813-814)
                    uint8_t *o = image + r * image_pitch + (int64_t)c * 3;
                    const uint8_t *s = colour + p * 3;
                    o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
                }
            }
        }
        uint8_t *mo = mask_img + r * mask_pitch + (int64_t)c * 3;
        mo[0] = m0; mo[1] = m1; mo[2] = m2;
    }
}

// --do_basic_infill (stereo_rerender.infill_using_normals, :155-240): one thread per hole pixel marches along the XY
// direction coded in the final mask image until it leaves the hole, then copies the colour found two / one / zero
// steps further on (the first that is inside the frame and not a hole).  Float32, one rounding per operation, like
// the NumPy code.  Only hole pixels are written and only non-hole pixels are read, so the image is updated in place.
template <bool FLOAT_NORMALS>
__global__ void __launch_bounds__(kThreads)
    normal_march_kernel(uint8_t *__restrict__ image, int64_t image_pitch, const uint8_t *__restrict__ hole, int64_t hole_pitch,
                        const uint8_t *__restrict__ mask_img, int64_t mask_pitch, const float *__restrict__ normal_map, int width, int height,
                        int max_steps) {
    const int64_t n = (int64_t)width * height;
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        const int y = (int)(p / width), x = (int)(p - (int64_t)y * width);
        if (!hole[y * hole_pitch + x]) continue;
        float dx, dy;
        if (FLOAT_NORMALS) {  // the function's own signature: a float normal map; (0, 1, 0) means "no normal here" (:176)
            const float *m = normal_map + p * 3;
            dx = m[0];
            dy = m[1];
            if (dx == 0.0f && dy == 1.0f && m[2] == 0.0f) continue;
        } else {
            const uint8_t *m = mask_img + y * mask_pitch + (int64_t)x * 3;
            dx = __fsub_rn(__fmul_rn(__fdiv_rn((float)m[0], 255.0f), 2.0f), 1.0f);  // ((m / 255) * 2) - 1 (:808,811)
            dy = __fsub_rn(__fmul_rn(__fdiv_rn((float)m[1], 255.0f), 2.0f), 1.0f);
        }
        const float norm = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        if (!(norm > 1e-6f)) continue;
        dx = __fdiv_rn(dx, norm);
        dy = __fdiv_rn(dy, norm);
        const float fx = (float)x, fy = (float)y;
        auto at = [&](int t, int &xi, int &yi) {
            xi = __float2int_rn(__fadd_rn(fx, __fmul_rn(dx, (float)t)));  // rint: half to even
            yi = __float2int_rn(__fadd_rn(fy, __fmul_rn(dy, (float)t)));
            return xi >= 0 && xi < width && yi >= 0 && yi < height;
        };
        for (int t = 1; t <= max_steps; ++t) {
            int xi, yi;
            if (!at(t, xi, yi)) break;                       // left the frame: the ray dies, the pixel stays as it is
            if (hole[yi * hole_pitch + xi]) continue;
            for (int dt = 2; dt >= 0; --dt) {
                int x2, y2;
                if (at(t + dt, x2, y2) && !hole[y2 * hole_pitch + x2]) {
                    const uint8_t *s = image + y2 * image_pitch + (int64_t)x2 * 3;
                    uint8_t *o = image + y * image_pitch + (int64_t)x * 3;
                    o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
                    break;
                }
            }
            break;
        }
    }
}

static int grid_for(int64_t work_items) {
    const int64_t blocks = (work_items + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

static Cam64 make_cam(const mdvt_source *s, const double *K) {
    Cam64 c{K[0], K[1], K[2], K[3], s->grid_sx, s->grid_sy, !(s->grid_sx == 1.0f && s->grid_sy == 1.0f)};
    return c;
}

}  // namespace mdvt

using namespace mdvt;

extern "C" int mdvt_edge_vertices(const void *depth_src, const mdvt_source *src, const double *K_host, double angle_threshold_deg,
                                  uint8_t *cell_flags_scratch, uint8_t *out_flags, double *out_normals, void *stream) {
    if (int rc = check_source(src)) return rc;
    MDVT_REQUIRE(src->width >= 2 && src->height >= 2, "the edge test needs at least a 2x2 grid");
    MDVT_REQUIRE(depth_src && K_host && cell_flags_scratch && out_flags, "NULL buffer");
    const Cam64 cam = make_cam(src, K_host);
    const double cos_limit = cos(angle_threshold_deg * (3.14159265358979323846 / 180.0));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t cells = (int64_t)(src->width - 1) * (src->height - 1), n = (int64_t)src->width * src->height;
#define CALL(D, B)                                                                                                                  \
    do {                                                                                                                            \
        edge_cells_kernel<D, B><<<grid_for(cells), kThreads, 0, st>>>(depth_src, src->width, src->height, src->dec_const, src->depth_scale, \
                                                                      cam, cos_limit, cell_flags_scratch);                          \
        edge_vertices_kernel<D, B><<<grid_for(n), kThreads, 0, st>>>(depth_src, src->width, src->height, src->dec_const, src->depth_scale,  \
                                                                     cam, cell_flags_scratch, out_flags, out_normals);              \
    } while (0)
    MDVT_DISPATCH_SOURCE(src->decoder, src->bit16, CALL);
#undef CALL
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_edge_vertices_xyz(const double *xyz, int width, int height, double angle_threshold_deg, uint8_t *cell_flags_scratch,
                                      uint8_t *out_flags, double *out_normals, void *stream) {
    MDVT_REQUIRE(width >= 2 && height >= 2, "the edge test needs at least a 2x2 grid");
    MDVT_REQUIRE(xyz && cell_flags_scratch && out_flags, "NULL buffer");
    const Cam64 cam{1.0, 1.0, 0.0, 0.0, 1.0f, 1.0f, 0};
    const double cos_limit = cos(angle_threshold_deg * (3.14159265358979323846 / 180.0));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t cells = (int64_t)(width - 1) * (height - 1), n = (int64_t)width * height;
    edge_cells_kernel<kSourceXYZ64, true><<<grid_for(cells), kThreads, 0, st>>>(xyz, width, height, 0.0f, 1.0f, cam, cos_limit, cell_flags_scratch);
    edge_vertices_kernel<kSourceXYZ64, true><<<grid_for(n), kThreads, 0, st>>>(xyz, width, height, 0.0f, 1.0f, cam, cell_flags_scratch, out_flags,
                                                                              out_normals);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_edge_splat(const void *depth_src, const mdvt_source *src, const double *K_host, const uint8_t *flags,
                               const double *pose_host, const double *K_render_host, int out_w, int out_h, uint64_t *zbuf, void *stream) {
    if (int rc = check_source(src)) return rc;
    MDVT_REQUIRE(depth_src && K_host && flags && pose_host && K_render_host && zbuf, "NULL buffer");
    MDVT_REQUIRE(out_w > 0 && out_h > 0, "bad output size %dx%d", out_w, out_h);
    const Cam64 cam = make_cam(src, K_host);
    Pose12d pose;
    for (int k = 0; k < 12; ++k) pose.m[k] = pose_host[k];
    const int64_t n = (int64_t)src->width * src->height;
    MDVT_REQUIRE(n <= 0xFFFFFFFFll, "source frame has more than 2^32 pixels");
    const double sqx = (double)(src->width - 1) / src->width, sqy = (double)(src->height - 1) / src->height;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(D, B)                                                                                                                 \
    edge_splat_kernel<D, B><<<grid_for(n), kThreads, 0, st>>>(depth_src, src->width, src->height, src->dec_const, src->depth_scale, cam, \
                                                              flags, pose, K_render_host[0], K_render_host[1], K_render_host[2],   \
                                                              K_render_host[3], sqx, sqy, out_w, out_h,                            \
                                                              reinterpret_cast<unsigned long long *>(zbuf))
    MDVT_DISPATCH_SOURCE(src->decoder, src->bit16, CALL);
#undef CALL
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_edge_resolve(uint64_t *zbuf, const void *depth_src, const mdvt_source *src, const double *K_host, const double *normals,
                                 const double *pose_host, const uint8_t *colour_rgb, const uint8_t *hole_mask, int64_t hole_pitch, int out_w,
                                 int out_h, uint32_t bg_rgb, int code_normals, uint8_t *image, int64_t image_pitch, uint8_t *mask_img,
                                 int64_t mask_pitch, void *stream) {
    if (int rc = check_source(src)) return rc;
    MDVT_REQUIRE(zbuf && depth_src && K_host && pose_host && colour_rgb && hole_mask && mask_img, "NULL buffer");
    MDVT_REQUIRE(!code_normals || normals, "normals are required to code them into the mask");
    MDVT_REQUIRE(out_w > 0 && out_h > 0, "bad output size %dx%d", out_w, out_h);
    MDVT_REQUIRE(hole_pitch >= out_w && mask_pitch >= (int64_t)out_w * 3 && (!image || image_pitch >= (int64_t)out_w * 3), "pitch too small");
    const Cam64 cam = make_cam(src, K_host);
    Pose12d pose;
    for (int k = 0; k < 12; ++k) pose.m[k] = pose_host[k];
    const double sqx = (double)(src->width - 1) / src->width, sqy = (double)(src->height - 1) / src->height;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(D, B)                                                                                                                  \
    edge_resolve_kernel<D, B><<<grid_for((int64_t)out_w * out_h), kThreads, 0, st>>>(                                               \
        reinterpret_cast<unsigned long long *>(zbuf), depth_src, src->width, src->height, src->dec_const, src->depth_scale, cam, normals, \
        pose, sqx, sqy, colour_rgb, hole_mask, hole_pitch, out_w, out_h, bg_rgb & 0xFFFFFF, code_normals, image, image_pitch, mask_img, \
        mask_pitch)
    MDVT_DISPATCH_SOURCE(src->decoder, src->bit16, CALL);
#undef CALL
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_normal_march_infill(uint8_t *image, int64_t image_pitch, const uint8_t *hole_mask, int64_t hole_pitch,
                                        const uint8_t *mask_img, int64_t mask_pitch, int width, int height, int max_steps, void *stream) {
    MDVT_REQUIRE(width > 0 && height > 0, "bad frame size %dx%d", width, height);
    MDVT_REQUIRE(image && hole_mask && mask_img, "NULL buffer");
    MDVT_REQUIRE(image_pitch >= (int64_t)width * 3 && mask_pitch >= (int64_t)width * 3 && hole_pitch >= width, "pitch too small");
    MDVT_REQUIRE(max_steps >= 0, "negative step count");
    normal_march_kernel<false><<<grid_for((int64_t)width * height), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        image, image_pitch, hole_mask, hole_pitch, mask_img, mask_pitch, nullptr, width, height, max_steps);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_normal_march_infill_f32(uint8_t *image, int64_t image_pitch, const uint8_t *hole_mask, int64_t hole_pitch,
                                            const float *normal_map, int width, int height, int max_steps, void *stream) {
    MDVT_REQUIRE(width > 0 && height > 0, "bad frame size %dx%d", width, height);
    MDVT_REQUIRE(image && hole_mask && normal_map, "NULL buffer");
    MDVT_REQUIRE(image_pitch >= (int64_t)width * 3 && hole_pitch >= width, "pitch too small");
    MDVT_REQUIRE(max_steps >= 0, "negative step count");
    normal_march_kernel<true><<<grid_for((int64_t)width * height), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        image, image_pitch, hole_mask, hole_pitch, nullptr, 0, normal_map, width, height, max_steps);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}
