// cv2.remap(INTER_LINEAR, BORDER_CONSTANT) on u8x3 images with float32 coordinate maps, bit-exact with OpenCV's
// fixed-point path: coordinates are quantised to 1/32 pixel (cvRound(map * 32), half to even), the four bilinear
// weights are the integers wx * wy * 32 (sum 2^15), the result is (sum(v * w) + 2^14) >> 15, and a tap that falls
// outside the source contributes the border value.  Used for the VR180 equirectangular output
// (stereo_rerender.convert_to_equirectangular, stereo_rerender.py:25-86,914-916).
#include "mdvt_common.cuh"

namespace mdvt {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
    remap_bilinear_u8x3_kernel(const uint8_t *__restrict__ src, int src_w, int src_h, int64_t src_pitch, const float *__restrict__ map_x,
                               const float *__restrict__ map_y, int dst_w, int dst_h, uint32_t border_rgb, uint8_t *__restrict__ dst,
                               int64_t dst_pitch) {
    const int64_t n = (int64_t)dst_w * dst_h;
    for (int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x; t < n; t += (int64_t)gridDim.x * kThreads) {
        const int r = (int)(t / dst_w), c = (int)(t - (int64_t)r * dst_w);
        // cvRound: round half to even; clamp first so that wild coordinates stay far outside instead of overflowing
        const float mx = fminf(fmaxf(__ldg(map_x + t), -1.0e6f), 1.0e6f), my = fminf(fmaxf(__ldg(map_y + t), -1.0e6f), 1.0e6f);
        const int sx = __float2int_rn(__fmul_rn(mx, 32.0f)), sy = __float2int_rn(__fmul_rn(my, 32.0f));
        const int ix = sx >> 5, iy = sy >> 5, fx = sx & 31, fy = sy & 31;
        int acc[3] = {1 << 14, 1 << 14, 1 << 14};
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int w = (dx ? fx : 32 - fx) * (dy ? fy : 32 - fy) * 32;
                if (w == 0) continue;
                const int x = ix + dx, y = iy + dy;
                uint32_t v0, v1, v2;
                if (x >= 0 && x < src_w && y >= 0 && y < src_h) {
                    const uint8_t *p = src + y * src_pitch + (int64_t)x * 3;
                    v0 = __ldg(p); v1 = __ldg(p + 1); v2 = __ldg(p + 2);
                } else {
                    v0 = border_rgb & 0xFF; v1 = (border_rgb >> 8) & 0xFF; v2 = (border_rgb >> 16) & 0xFF;
                }
                acc[0] += (int)v0 * w; acc[1] += (int)v1 * w; acc[2] += (int)v2 * w;
            }
        }
        uint8_t *o = dst + r * dst_pitch + (int64_t)c * 3;
        o[0] = (uint8_t)(acc[0] >> 15); o[1] = (uint8_t)(acc[1] >> 15); o[2] = (uint8_t)(acc[2] >> 15);
    }
}

}  // namespace mdvt

using namespace mdvt;

extern "C" int mdvt_remap_bilinear_u8x3(const uint8_t *src, int src_w, int src_h, int64_t src_pitch, const float *map_x, const float *map_y,
                                        int dst_w, int dst_h, uint32_t border_rgb, uint8_t *dst, int64_t dst_pitch, void *stream) {
    MDVT_REQUIRE(src_w > 0 && src_h > 0 && dst_w > 0 && dst_h > 0, "bad image size");
    MDVT_REQUIRE(src && map_x && map_y && dst, "NULL buffer");
    MDVT_REQUIRE(src_pitch >= (int64_t)src_w * 3 && dst_pitch >= (int64_t)dst_w * 3, "pitch too small");
    const int64_t n = (int64_t)dst_w * dst_h;
    const int64_t blocks = (n + kThreads - 1) / kThreads, cap = (int64_t)sm_count() * 8;
    remap_bilinear_u8x3_kernel<<<(int)(blocks < cap ? blocks : cap), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        src, src_w, src_h, src_pitch, map_x, map_y, dst_w, dst_h, border_rgb & 0xFFFFFF, dst, dst_pitch);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}
