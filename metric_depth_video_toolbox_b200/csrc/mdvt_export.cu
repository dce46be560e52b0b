// Depth export formats written by the scripts besides the RGB-coded wire format (all pointwise, HBM-bound):
//   * 16-bit / 8-bit grey video frames (convert_metric_depth_video_to_other_format.py:752-760)
//   * Touchly reverse-depth planes (stereo_rerender.py:548-552,687-690,826-829)
#include "mdvt_common.cuh"

namespace mdvt {

constexpr int kThreads = 256;

static int grid_for(int64_t work_items) {
    const int64_t blocks = (work_items + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

// out = astype(rint(depth * factor)): NumPy multiplies the float32 plane by the Python scalar in float32, rounds half
// to even, and the integer cast goes through int32 (values past the range wrap, as NumPy does on x86).
template <int DECODER, bool BIT16, typename OUT, int CHANNELS>
__global__ void __launch_bounds__(kThreads)
    depth_to_grey_kernel(const void *__restrict__ src, int64_t n, float dec_const, float factor, OUT *__restrict__ out) {
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        const float v = rintf(__fmul_rn(source_depth<DECODER, BIT16>(src, p, dec_const), factor));
        const OUT q = (OUT)(int32_t)v;
#pragma unroll
        for (int c = 0; c < CHANNELS; ++c) out[p * CHANNELS + c] = q;
    }
}

// 255 - rint(max(0, min(depth, tmax) - tmin) * fl32(255 / (tmax - tmin))) replicated to 3 channels; with
// zero_is_far a quantised 0 (nothing rendered there) is moved to the far end first (:688,:827).
template <int DECODER, bool BIT16>
__global__ void __launch_bounds__(kThreads)
    touchly_depth_kernel(const void *__restrict__ src, int64_t n, float dec_const, float depth_scale, float tmin, float tmax, float gain,
                         int zero_is_far, uint8_t *__restrict__ out, int64_t width, int64_t out_pitch) {
    for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kThreads) {
        const float d = __fmul_rn(source_depth<DECODER, BIT16>(src, p, dec_const), depth_scale);
        const float clipped = fmaxf(0.0f, __fsub_rn(fminf(d, tmax), tmin));
        uint32_t q = (uint32_t)(int32_t)rintf(__fmul_rn(clipped, gain)) & 0xFFu;
        if (zero_is_far && q == 0) q = 255;
        q = 255u - q;
        const int64_t row = p / width, col = p - row * width;
        uint8_t *o = out + row * out_pitch + col * 3;
        o[0] = o[1] = o[2] = (uint8_t)q;
    }
}

// Hole mask u8 {0, non-zero} -> 1 bit per pixel, most significant bit first (np.packbits order): what crosses PCIe when
// the host API is asked for packed masks (a 1080p stereo mask shrinks from 4.1 MB to 0.5 MB per frame).  16 pixels per
// thread: one 16-byte load, the top bit of every byte (0xFF / 0x00 masks; any non-zero byte counts) gathered by a
// multiply, one 2-byte store.
__global__ void __launch_bounds__(kThreads) pack_mask_bits_kernel(const uint8_t *__restrict__ mask, int64_t n_groups16, int64_t n_pixels,
                                                                  uint8_t *__restrict__ bits) {
    for (int64_t g = blockIdx.x * (int64_t)kThreads + threadIdx.x; g < n_groups16; g += (int64_t)gridDim.x * kThreads) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if ((g + 1) * 16 <= n_pixels) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(mask) + g);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
            for (int k = 0; k < 16 && g * 16 + k < n_pixels; ++k) w[k >> 2] |= (uint32_t)mask[g * 16 + k] << (8 * (k & 3));
        }
        uint32_t out = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            // non-zero byte -> its top bit set: (x | (x + 0x7F7F7F7F per byte without carry across bytes)) & 0x80
            const uint32_t x = w[q];
            const uint32_t nz = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
            // bytes 0..3 (pixels 4q .. 4q+3) -> bits 3..0 of a nibble, first pixel most significant
            const uint32_t nib = ((nz >> 7) * 0x08040201u >> 24) & 0xFu;   // byte0 -> bit 3, byte1 -> bit 2, byte2 -> bit 1, byte3 -> bit 0
            out |= nib << (4 * ((q & 1) ? 0 : 1) + 8 * (q >> 1));
        }
        const int64_t b = g * 2;
        const int64_t n_bytes = (n_pixels + 7) >> 3;
        if (b + 1 < n_bytes) *reinterpret_cast<uint16_t *>(bits + b) = (uint16_t)out;
        else if (b < n_bytes) bits[b] = (uint8_t)out;
    }
}

}  // namespace mdvt

using namespace mdvt;

extern "C" int mdvt_pack_mask_bits(const uint8_t *mask, int64_t n_pixels, uint8_t *bits, void *stream) {
    MDVT_REQUIRE(n_pixels >= 0, "negative pixel count");
    if (n_pixels == 0) return MDVT_OK;
    MDVT_REQUIRE(mask && bits, "NULL buffer");
    MDVT_REQUIRE((reinterpret_cast<uintptr_t>(mask) & 15) == 0 && (reinterpret_cast<uintptr_t>(bits) & 1) == 0, "mask must be 16-byte, bits 2-byte aligned");
    const int64_t groups = (n_pixels + 15) / 16;
    pack_mask_bits_kernel<<<grid_for(groups), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(mask, groups, n_pixels, bits);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_depth_to_grey(const void *depth_src, int64_t n_pixels, int decoder, int bit16, float dec_const, float factor,
                                  int out_bits, int channels, void *out, void *stream) {
    MDVT_REQUIRE(n_pixels >= 0, "negative pixel count");
    if (int rc = check_decoder(decoder, bit16, true)) return rc;
    MDVT_REQUIRE((out_bits == 8 && (channels == 1 || channels == 3)) || (out_bits == 16 && channels == 1),
                 "supported outputs: 8-bit x1, 8-bit x3, 16-bit x1");
    if (n_pixels == 0) return MDVT_OK;
    MDVT_REQUIRE(depth_src && out, "NULL buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(n_pixels);
#define CALL(D, B)                                                                                                              \
    do {                                                                                                                        \
        if (out_bits == 16)                                                                                                     \
            depth_to_grey_kernel<D, B, uint16_t, 1><<<grid, kThreads, 0, st>>>(depth_src, n_pixels, dec_const, factor, (uint16_t *)out); \
        else if (channels == 3)                                                                                                 \
            depth_to_grey_kernel<D, B, uint8_t, 3><<<grid, kThreads, 0, st>>>(depth_src, n_pixels, dec_const, factor, (uint8_t *)out);   \
        else                                                                                                                    \
            depth_to_grey_kernel<D, B, uint8_t, 1><<<grid, kThreads, 0, st>>>(depth_src, n_pixels, dec_const, factor, (uint8_t *)out);   \
    } while (0)
    MDVT_DISPATCH_SOURCE(decoder, bit16, CALL);
#undef CALL
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_touchly_depth(const void *depth_src, int width, int height, int decoder, int bit16, float dec_const, float depth_scale,
                                  float touchly_min, float touchly_max, float gain, int zero_is_far, uint8_t *out_rgb, int64_t out_pitch,
                                  void *stream) {
    MDVT_REQUIRE(width > 0 && height > 0, "bad frame size %dx%d", width, height);
    if (int rc = check_decoder(decoder, bit16, true)) return rc;
    MDVT_REQUIRE(touchly_max > touchly_min, "touchly_max_depth must exceed touchly_min_depth");
    MDVT_REQUIRE(depth_src && out_rgb, "NULL buffer");
    MDVT_REQUIRE(out_pitch >= (int64_t)width * 3, "out_pitch %lld too small", (long long)out_pitch);
    const int64_t n = (int64_t)width * height;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(D, B)                                                                                                                 \
    touchly_depth_kernel<D, B><<<grid_for(n), kThreads, 0, st>>>(depth_src, n, dec_const, depth_scale, touchly_min, touchly_max, gain, \
                                                                 zero_is_far, out_rgb, width, out_pitch)
    MDVT_DISPATCH_SOURCE(decoder, bit16, CALL);
#undef CALL
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}
