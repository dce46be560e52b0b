// FFV1 (version 3, Golomb-Rice, RGB) encoder on the device: one thread per slice, up to 1024 slices per frame and a batch
// of frames per launch, slices compacted into ready-to-mux packets on the device.  The result videos of the reference
// all go through cv2.VideoWriter(fourcc "FFV1") (stereo_rerender.py:420-442,941; depth_frames_helper.py:125-161;
// 3d_view_depthfile.py:118-127); on the host that entropy coder costs ~0.45 core-seconds per 3840x1080 frame and caps
// the files-in/files-out rate.  Streams written here are decoded bit-exactly by libavcodec.
//
// Host side (no device needed): the range coder of rangecoder.c for the configuration record (ffv1enc.c
// write_extradata) and the per-slice headers (encode_slice_header); tests pin both byte-for-byte against OpenCV's files.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mdvt_common.cuh"

#define MDVT_FFV1_HD __host__ __device__
#ifdef __CUDA_ARCH__
#define MDVT_FFV1_LD(p) __ldg(p)
#endif
#include "mdvt_ffv1_slice.h"

namespace mdvt {
namespace {

// ---- host: range coder (rangecoder.c), symbols (ffv1enc.c put_symbol_inline), CRC ------------------------------------
struct RangeCoder {
    uint8_t zero_state[256], one_state[256];
    uint32_t low = 0, range = 0xFF00;
    int outstanding_count = 0, outstanding_byte = -1;
    std::vector<uint8_t> out;

    RangeCoder() {   // ff_build_rac_states(c, 0.05 * (1LL << 32), 256 - 8)
        const int64_t one = 1LL << 32, factor = (int64_t)(0.05 * (double)(1LL << 32));
        const int max_p = 256 - 8;
        memset(zero_state, 0, sizeof zero_state);
        memset(one_state, 0, sizeof one_state);
        int last_p8 = 0;
        int64_t p = one / 2;
        for (int i = 0; i < 128; ++i) {
            int p8 = (int)((256 * p + one / 2) >> 32);
            if (p8 <= last_p8) p8 = last_p8 + 1;
            if (last_p8 && last_p8 < 256 && p8 <= max_p) one_state[last_p8] = (uint8_t)p8;
            p += ((one - p) * factor + one / 2) >> 32;
            last_p8 = p8;
        }
        for (int i = 256 - max_p; i <= max_p; ++i) {
            if (one_state[i]) continue;
            p = ((int64_t)i * one + 128) >> 8;
            p += ((one - p) * factor + one / 2) >> 32;
            int p8 = (int)((256 * p + one / 2) >> 32);
            if (p8 <= i) p8 = i + 1;
            if (p8 > max_p) p8 = max_p;
            one_state[i] = (uint8_t)p8;
        }
        for (int i = 1; i < 255; ++i) zero_state[i] = (uint8_t)(256 - one_state[256 - i]);
    }
    void renorm() {
        if (outstanding_byte < 0) {
            outstanding_byte = (int)(low >> 8);
        } else if (low <= 0xFF00) {
            out.push_back((uint8_t)outstanding_byte);
            out.insert(out.end(), outstanding_count, 0xFF);
            outstanding_count = 0;
            outstanding_byte = (int)(low >> 8);
        } else if (low >= 0x10000) {
            out.push_back((uint8_t)(outstanding_byte + 1));
            out.insert(out.end(), outstanding_count, 0x00);
            outstanding_count = 0;
            outstanding_byte = (int)((low >> 8) & 0xFF);
        } else {
            ++outstanding_count;
        }
        low = (low & 0xFF) << 8;
        range <<= 8;
    }
    void put_rac(uint8_t *state, int bit) {
        const uint32_t range1 = (range * *state) >> 8;
        if (!bit) {
            range -= range1;
            *state = zero_state[*state];
        } else {
            low += range - range1;
            range = range1;
            *state = one_state[*state];
        }
        while (range < 0x100) renorm();
    }
    void put_symbol(uint8_t *state, int v, bool is_signed) {
        if (!v) {
            put_rac(state + 0, 1);
            return;
        }
        const int a = v < 0 ? -v : v;
        int e = 0;
        while ((a >> (e + 1)) != 0) ++e;
        put_rac(state + 0, 0);
        for (int i = 0; i < e; ++i) put_rac(state + 1 + (i < 9 ? i : 9), 1);
        put_rac(state + 1 + (e < 9 ? e : 9), 0);
        for (int i = e - 1; i >= 0; --i) put_rac(state + 22 + (i < 9 ? i : 9), (a >> i) & 1);
        if (is_signed) put_rac(state + 11 + (e < 10 ? e : 10), v < 0);
    }
    void terminate(bool sentinel) {   // ff_rac_terminate
        if (sentinel) {
            uint8_t s = 129;
            put_rac(&s, 0);
        }
        range = 0xFF;
        low += 0xFF;
        renorm();
        range = 0xFF;
        renorm();
    }
};

// The reading side of the same coder (rangecoder.h get_rac / ffv1.h get_symbol): only the leading fields of a
// configuration record are read here.
struct RangeReader {
    RangeCoder tables;   // state transition tables
    const uint8_t *buf;
    int pos = 2, end;
    uint32_t low, range = 0xFF00;
    bool bad = false;

    RangeReader(const uint8_t *b, int n) : buf(b), end(n) {
        low = n >= 2 ? ((uint32_t)b[0] << 8) | b[1] : 0;
        if (n < 2) bad = true;
        if (low >= 0xFF00) {
            low = 0xFF00;
            end = pos;
        }
    }
    void refill() {
        if (range < 0x100) {
            range <<= 8;
            low <<= 8;
            if (pos < end) low += buf[pos++];
        }
    }
    int get_rac(uint8_t *state) {
        const uint32_t range1 = (range * *state) >> 8;
        range -= range1;
        if (low < range) {
            *state = tables.zero_state[*state];
            refill();
            return 0;
        }
        low -= range;
        *state = tables.one_state[*state];
        range = range1;
        refill();
        return 1;
    }
    int get_symbol(uint8_t *state) {   // unsigned symbols only
        if (get_rac(state + 0)) return 0;
        int e = 0;
        while (get_rac(state + 1 + (e < 9 ? e : 9))) {
            if (++e > 31) {
                bad = true;
                return 0;
            }
        }
        int a = 1;
        for (int i = e - 1; i >= 0; --i) a += a + get_rac(state + 22 + (i < 9 ? i : 9));
        return a;
    }
};

struct CrcTable {
    uint32_t t[256];
    CrcTable() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i << 24;
            for (int k = 0; k < 8; ++k) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : c << 1;
            t[i] = c;
        }
    }
};
const CrcTable &crc_table_host() {
    static const CrcTable table;
    return table;
}

uint32_t crc32_mpeg(const uint8_t *p, size_t n) {
    const uint32_t *t = crc_table_host().t;
    uint32_t c = 0;
    for (size_t i = 0; i < n; ++i) c = (c << 8) ^ t[(c >> 24) ^ p[i]];
    return c;
}

void write_run_table(RangeCoder &rc, const int *runs, int n) {   // write_quant_table: run lengths of q[0..127]
    uint8_t state[32];
    memset(state, 128, sizeof state);
    for (int i = 0; i < n; ++i) rc.put_symbol(state, runs[i] - 1, false);
}

int check_stream(int width, int height, int nh, int nv, int alpha) {
    MDVT_REQUIRE(width > 0 && height > 0 && width <= 65535 && height <= 65535, "bad frame size %dx%d", width, height);
    MDVT_REQUIRE(nh >= 1 && nv >= 1 && nh <= width && nv <= height, "bad slice grid %dx%d for a %dx%d frame", nh, nv, width, height);
    MDVT_REQUIRE(nh * nv <= 1024, "%d x %d slices: FFV1 allows at most 1024 per frame", nh, nv);   // ffv1.h MAX_SLICES
    MDVT_REQUIRE(alpha == 0 || alpha == 1, "alpha must be 0 or 1");
    return MDVT_OK;
}

int check_model(int context_model) {
    MDVT_REQUIRE(context_model >= 0 && context_model <= 2, "context_model must be 0 (libavcodec's 666 contexts), 1 (63 contexts) or 2 (14 contexts)");
    return MDVT_OK;
}

// Four CRC tables for word-at-a-time updates: g_crc_table[k * 256 + i] = CRC of byte i followed by k zero bytes.
__device__ uint32_t g_crc_table[1024];

__device__ __forceinline__ uint32_t crc_of_byte(uint32_t b) {
    uint32_t c = b << 24;
    for (int k = 0; k < 8; ++k) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : c << 1;
    return c;
}

__global__ void ffv1_crc_table_kernel() {
    uint32_t c = crc_of_byte(threadIdx.x);
    g_crc_table[threadIdx.x] = c;
    for (int k = 1; k < 4; ++k) {
        c = (c << 8) ^ crc_of_byte(c >> 24);
        g_crc_table[k * 256 + threadIdx.x] = c;
    }
}

// One thread per slice.  Consecutive threads take horizontally adjacent slices of one frame, so a warp walks neighbouring
// row segments (shared 128-byte lines); each thread owns a contiguous state block and a contiguous output range.
__global__ void __launch_bounds__(64) ffv1_encode_kernel(const uint8_t *__restrict__ frames, int64_t frame_stride, int64_t row_pitch,
                                                         int n_frames, int width, int height, int nh, int nv, int n_planes, int ib,
                                                         int ir, int model, const uint8_t *__restrict__ headers,
                                                         const int32_t *__restrict__ header_len, mdvt_ffv1::VlcState *states,
                                                         uint8_t *out, int64_t capacity, int32_t *sizes) {
    __shared__ uint32_t crc_s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) crc_s[i] = g_crc_table[i];
    __syncthreads();
    const int per_frame = nh * nv;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_frames * per_frame) return;
    const int f = (int)(t / per_frame), si = (int)(t - (int64_t)f * per_frame);
    const int sy = si / nh, sx = si - sy * nh;
    const int x0 = (int)((int64_t)sx * width / nh), x1 = (int)((int64_t)(sx + 1) * width / nh);
    const int y0 = (int)((int64_t)sy * height / nv), y1 = (int)((int64_t)(sy + 1) * height / nv);
    mdvt_ffv1::SliceJob job;
    job.frame = frames + f * frame_stride + y0 * row_pitch + 3 * (int64_t)x0;
    job.row_pitch = row_pitch;
    job.w = x1 - x0;
    job.h = y1 - y0;
    job.n_planes = n_planes;
    job.ib = ib;
    job.ir = ir;
    job.header = headers + si * mdvt_ffv1::kHeaderStride;
    job.header_len = header_len[si];
    job.states = states + t * (int64_t)((n_planes > 3 ? 3 : 2) * mdvt_ffv1::contexts_of(model));
    job.out = out + t * capacity;
    job.crc_table = crc_s;
    sizes[t] = (int32_t)(model == 2 ? mdvt_ffv1::encode_slice<2>(job) : (model ? mdvt_ffv1::encode_slice<1>(job) : mdvt_ffv1::encode_slice<0>(job)));
}

// Context model 2 without the alpha plane: the 2 x 14 states of a slice (224 bytes) live in SHARED memory, context-major and
// interleaved over the CTA's threads (state of context c of thread t at [c * TPB + t]: the threads of a warp that sit in the
// same context hit consecutive banks).  14.3 KB per 64 slices: the register file, not the shared memory, bounds the resident
// warps, and a state access costs a shared-memory round trip instead of an L1 / L2 one (with 1 KB per slice, model 1, the same
// layout leaves 6 warps per SM and buys nothing: profiles/r02_ffv1_state_experiments.txt).
template <int TPB>
__global__ void __launch_bounds__(TPB) ffv1_encode_tiny_kernel(const uint8_t *__restrict__ frames, int64_t frame_stride, int64_t row_pitch,
                                                               int n_frames, int width, int height, int nh, int nv, int ib, int ir,
                                                               const uint8_t *__restrict__ headers, const int32_t *__restrict__ header_len,
                                                               uint8_t *out, int64_t capacity, int32_t *sizes) {
    __shared__ uint32_t crc_s[1024];
    __shared__ mdvt_ffv1::VlcState state_s[2 * mdvt_ffv1::kContextsTiny * TPB];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) crc_s[i] = g_crc_table[i];
    __syncthreads();
    const int per_frame = nh * nv;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_frames * per_frame) return;
    const int f = (int)(t / per_frame), si = (int)(t - (int64_t)f * per_frame);
    const int sy = si / nh, sx = si - sy * nh;
    const int x0 = (int)((int64_t)sx * width / nh), x1 = (int)((int64_t)(sx + 1) * width / nh);
    const int y0 = (int)((int64_t)sy * height / nv), y1 = (int)((int64_t)(sy + 1) * height / nv);
    mdvt_ffv1::SliceJob job;
    job.frame = frames + f * frame_stride + y0 * row_pitch + 3 * (int64_t)x0;
    job.row_pitch = row_pitch;
    job.w = x1 - x0;
    job.h = y1 - y0;
    job.n_planes = 3;
    job.ib = ib;
    job.ir = ir;
    job.header = headers + si * mdvt_ffv1::kHeaderStride;
    job.header_len = header_len[si];
    job.states = state_s + threadIdx.x;
    job.out = out + t * capacity;
    job.crc_table = crc_s;
    sizes[t] = (int32_t)mdvt_ffv1::encode_slice<2, TPB>(job);
}

// ---- decoder ----------------------------------------------------------------------------------------------------------
// Slice boundaries of every packet: the footers are walked backwards from the end of the packet (ffv1dec.c decode_frame:
// each footer holds the size of its slice).  One thread per frame.  status[f] = 0, or -2 when the sizes do not add up.
__global__ void ffv1_index_kernel(const uint8_t *__restrict__ packets, const int64_t *__restrict__ packet_offsets, int n_frames,
                                  int per_frame, int64_t *__restrict__ slice_offsets, int32_t *__restrict__ status) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const int64_t begin = packet_offsets[f];
    int64_t p = packet_offsets[f + 1];
    int code = 0;
    for (int s = per_frame - 1; s >= 0; --s) {
        int64_t size = -1;
        if (p - begin >= mdvt_ffv1::kFooterBytes) {
            const uint8_t *foot = packets + p - mdvt_ffv1::kFooterBytes;
            size = (((int64_t)foot[0] << 16) | ((int64_t)foot[1] << 8) | (int64_t)foot[2]) + mdvt_ffv1::kFooterBytes;
        }
        if (size < 0 || p - size < begin) {   // malformed: give the remaining slices empty ranges
            code = -2;
            size = 0;
        }
        p -= size;
        slice_offsets[(int64_t)f * per_frame + s] = p;
    }
    if (p != begin) code = -2;
    status[f] = code;
}

__global__ void __launch_bounds__(64, 12) ffv1_decode_kernel(const uint8_t *__restrict__ packets, const int64_t *__restrict__ packet_offsets,
                                                         const int64_t *__restrict__ slice_offsets, int n_frames, int width, int height,
                                                         int nh, int nv, int n_planes, int ib, int ir, int model,
                                                         const uint8_t *__restrict__ headers, const int32_t *__restrict__ header_len,
                                                         mdvt_ffv1::VlcState *states, uint8_t *frames, int64_t frame_stride,
                                                         int64_t row_pitch, int32_t *status) {
    __shared__ uint32_t crc_s[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) crc_s[i] = g_crc_table[i];
    __syncthreads();
    const int per_frame = nh * nv;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_frames * per_frame) return;
    const int f = (int)(t / per_frame), si = (int)(t - (int64_t)f * per_frame);
    if (status[f] == -2) return;   // the index kernel could not split this packet
    const int sy = si / nh, sx = si - sy * nh;
    const int x0 = (int)((int64_t)sx * width / nh), x1 = (int)((int64_t)(sx + 1) * width / nh);
    const int y0 = (int)((int64_t)sy * height / nv), y1 = (int)((int64_t)(sy + 1) * height / nv);
    const int64_t begin = slice_offsets[t];
    const int64_t end = si + 1 < per_frame ? slice_offsets[t + 1] : packet_offsets[f + 1];
    mdvt_ffv1::SliceInput in;
    in.data = packets + begin;
    in.size = (uint32_t)(end - begin);
    in.header = headers + si * mdvt_ffv1::kHeaderStride;
    in.header_len = header_len[si];
    in.frame = frames + f * frame_stride + y0 * row_pitch + 3 * (int64_t)x0;
    in.row_pitch = row_pitch;
    in.w = x1 - x0;
    in.h = y1 - y0;
    in.n_planes = n_planes;
    in.ib = ib;
    in.ir = ir;
    in.states = states + t * (int64_t)((n_planes > 3 ? 3 : 2) * mdvt_ffv1::contexts_of(model));
    in.crc_table = crc_s;
    const int code = model == 2 ? mdvt_ffv1::decode_slice<2>(in) : (model ? mdvt_ffv1::decode_slice<1>(in) : mdvt_ffv1::decode_slice<0>(in));
    if (code < 0) atomicMin(&status[f], code - 2);   // -3: foreign slice header, -4: slice size, -5: bit stream overrun, -6: CRC
}

// The decoder for context model 2 without the alpha plane, coder states in shared memory (see ffv1_encode_tiny_kernel).
template <int TPB>
__global__ void __launch_bounds__(TPB, 12) ffv1_decode_tiny_kernel(const uint8_t *__restrict__ packets, const int64_t *__restrict__ packet_offsets,
                                                                   const int64_t *__restrict__ slice_offsets, int n_frames, int width, int height,
                                                                   int nh, int nv, int ib, int ir, const uint8_t *__restrict__ headers,
                                                                   const int32_t *__restrict__ header_len, uint8_t *frames, int64_t frame_stride,
                                                                   int64_t row_pitch, int32_t *status) {
    __shared__ uint32_t crc_s[256];
    __shared__ mdvt_ffv1::VlcState state_s[2 * mdvt_ffv1::kContextsTiny * TPB];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) crc_s[i] = g_crc_table[i];
    __syncthreads();
    const int per_frame = nh * nv;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_frames * per_frame) return;
    const int f = (int)(t / per_frame), si = (int)(t - (int64_t)f * per_frame);
    if (status[f] == -2) return;   // the index kernel could not split this packet
    const int sy = si / nh, sx = si - sy * nh;
    const int x0 = (int)((int64_t)sx * width / nh), x1 = (int)((int64_t)(sx + 1) * width / nh);
    const int y0 = (int)((int64_t)sy * height / nv), y1 = (int)((int64_t)(sy + 1) * height / nv);
    const int64_t begin = slice_offsets[t];
    const int64_t end = si + 1 < per_frame ? slice_offsets[t + 1] : packet_offsets[f + 1];
    mdvt_ffv1::SliceInput in;
    in.data = packets + begin;
    in.size = (uint32_t)(end - begin);
    in.header = headers + si * mdvt_ffv1::kHeaderStride;
    in.header_len = header_len[si];
    in.frame = frames + f * frame_stride + y0 * row_pitch + 3 * (int64_t)x0;
    in.row_pitch = row_pitch;
    in.w = x1 - x0;
    in.h = y1 - y0;
    in.n_planes = 3;
    in.ib = ib;
    in.ir = ir;
    in.states = state_s + threadIdx.x;
    in.crc_table = crc_s;
    const int code = mdvt_ffv1::decode_slice<2, TPB>(in);
    if (code < 0) atomicMin(&status[f], code - 2);   // -3: foreign slice header, -4: slice size, -5: bit stream overrun, -6: CRC
}

// Packet layout: offsets[f * S + s] = first byte of slice s of frame f in the packed stream, offsets[n * S] = total.
// One CTA; frames in sequence, a block-wide scan over the (<= 1024) slices of each.
__global__ void __launch_bounds__(1024) ffv1_offsets_kernel(const int32_t *__restrict__ sizes, int n_frames, int per_frame,
                                                            int64_t *__restrict__ offsets) {
    __shared__ int32_t warp_sum[32];
    __shared__ int64_t base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int f = 0; f < n_frames; ++f) {
        const int32_t mine = threadIdx.x < per_frame ? sizes[(int64_t)f * per_frame + threadIdx.x] : 0;
        int32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int32_t w = warp_sum[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t up = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= d) wi += up;
            }
            warp_sum[lane] = wi - w;   // exclusive
        }
        __syncthreads();
        const int64_t base = base_s;
        const int64_t excl = base + warp_sum[warp] + (incl - mine);
        if (threadIdx.x < per_frame) offsets[(int64_t)f * per_frame + threadIdx.x] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) base_s = base + warp_sum[31] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[(int64_t)n_frames * per_frame] = base_s;
}

// One CTA per slice: copies its bytes to their place in the packed stream.
__global__ void __launch_bounds__(128) ffv1_pack_kernel(const uint8_t *__restrict__ slices, int64_t capacity,
                                                        const int32_t *__restrict__ sizes, const int64_t *__restrict__ offsets,
                                                        uint8_t *__restrict__ packed) {
    const int64_t s = blockIdx.x;
    const uint8_t *src = slices + s * capacity;   // 16-byte aligned (capacity is a multiple of 16)
    uint8_t *dst = packed + offsets[s];
    const int n = sizes[s];
    const int head = (int)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
    for (int i = threadIdx.x; i < n && i < head; i += blockDim.x) dst[i] = src[i];
    // body: aligned 16-byte stores, source read as bytes through the read-only path (misaligned relative to dst)
    const int body = n > head ? (n - head) & ~15 : 0;
    for (int i = threadIdx.x * 16; i < body; i += blockDim.x * 16) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint8_t *p = src + head + i + 4 * k;
            w[k] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | ((uint32_t)__ldg(p + 3) << 24);
        }
        *reinterpret_cast<uint4 *>(dst + head + i) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (int i = head + body + threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace
}  // namespace mdvt

extern "C" int64_t mdvt_ffv1_slice_capacity(int width, int height, int nh, int nv, int alpha) {
    if (mdvt::check_stream(width, height, nh, nv, alpha) != MDVT_OK) return -1;
    const int w = (width + nh - 1) / nh + 1, h = (height + nv - 1) / nv + 1;
    return mdvt_ffv1::slice_capacity(w, h, 3 + alpha);
}

extern "C" int64_t mdvt_ffv1_state_bytes(int n_frames, int nh, int nv, int alpha, int context_model) {
    if (n_frames < 0 || nh < 1 || nv < 1 || nh * nv > 1024 || context_model < 0 || context_model > 2) return -1;
    return (int64_t)n_frames * nh * nv * (alpha ? 3 : 2) * mdvt_ffv1::contexts_of(context_model) * (int64_t)sizeof(mdvt_ffv1::VlcState);
}

extern "C" int mdvt_ffv1_stream_setup(int width, int height, int nh, int nv, int alpha, int context_model, uint8_t *config_host,
                                      int config_capacity, int *config_len, uint8_t *headers_host, int32_t *header_len_host) {
    using mdvt::RangeCoder;
    if (int rc = mdvt::check_stream(width, height, nh, nv, alpha)) return rc;
    if (int rc = mdvt::check_model(context_model)) return rc;
    MDVT_REQUIRE(config_host && config_len && headers_host && header_len_host, "NULL output");
    static const int q11[] = {1, 1, 3, 7, 23, 93}, q5[] = {1, 3, 124}, q3[] = {1, 127}, q0[] = {128};
    {   // ffv1enc.c write_extradata: version 3.4, Golomb-Rice, RGB, 8 bit; model 0: both of libavcodec's quant-table sets
        // (as its encoder writes them), model 1: one set, the 5-level table on three inputs, model 2: one set, the 3-level table
        RangeCoder rc;
        uint8_t st[32];
        memset(st, 128, sizeof st);
        rc.put_symbol(st, 3, false);          // version
        rc.put_symbol(st, 4, false);          // micro_version
        rc.put_symbol(st, 0, false);          // coder type: Golomb-Rice
        rc.put_symbol(st, 1, false);          // colourspace: RGB (JPEG2000 RCT)
        rc.put_symbol(st, 8, false);          // bits per raw sample
        rc.put_rac(st, 1);                    // chroma planes
        rc.put_symbol(st, 0, false);          // chroma shifts
        rc.put_symbol(st, 0, false);
        rc.put_rac(st, alpha);                // transparency
        rc.put_symbol(st, nh - 1, false);
        rc.put_symbol(st, nv - 1, false);
        if (context_model == 0) {
            rc.put_symbol(st, 2, false);      // quant table sets
            for (int set = 0; set < 2; ++set) {
                mdvt::write_run_table(rc, q11, 6);
                mdvt::write_run_table(rc, q11, 6);
                for (int k = 2; k < 5; ++k) {
                    if (set == 0 && k == 2) mdvt::write_run_table(rc, q11, 6);
                    else if (set == 0) mdvt::write_run_table(rc, q0, 1);
                    else mdvt::write_run_table(rc, q5, 3);
                }
            }
            rc.put_rac(st, 0);                // no coded initial states, per set
            rc.put_rac(st, 0);
        } else {
            rc.put_symbol(st, 1, false);
            for (int k = 0; k < 3; ++k) {
                if (context_model == 1) mdvt::write_run_table(rc, q5, 3);
                else mdvt::write_run_table(rc, q3, 2);
            }
            for (int k = 3; k < 5; ++k) mdvt::write_run_table(rc, q0, 1);
            rc.put_rac(st, 0);
        }
        rc.put_symbol(st, 1, false);          // ec: per-slice CRC
        rc.put_symbol(st, 0, false);          // intra flag as libavcodec writes it for gop_size > 1
        rc.terminate(false);
        const uint32_t crc = mdvt::crc32_mpeg(rc.out.data(), rc.out.size());
        for (int k = 3; k >= 0; --k) rc.out.push_back((uint8_t)(crc >> (8 * k)));
        MDVT_REQUIRE((int)rc.out.size() <= config_capacity, "configuration record needs %d bytes", (int)rc.out.size());
        memcpy(config_host, rc.out.data(), rc.out.size());
        *config_len = (int)rc.out.size();
    }
    for (int sy = 0; sy < nv; ++sy)
        for (int sx = 0; sx < nh; ++sx) {   // ffv1enc.c encode_slice_header; every frame is a key frame
            RangeCoder rc;
            const int si = sy * nh + sx;
            if (si == 0) {
                uint8_t key_state = 128;
                rc.put_rac(&key_state, 1);
            }
            uint8_t st[32];
            memset(st, 128, sizeof st);
            rc.put_symbol(st, sx, false);
            rc.put_symbol(st, sy, false);
            rc.put_symbol(st, 0, false);      // slice width - 1, height - 1 in grid units
            rc.put_symbol(st, 0, false);
            for (int k = 0; k < 2 + alpha; ++k) rc.put_symbol(st, 0, false);   // quant table set per plane context
            rc.put_symbol(st, 3, false);      // progressive
            rc.put_symbol(st, 0, false);      // sample aspect ratio 0/1, as OpenCV's files have it
            rc.put_symbol(st, 1, false);
            rc.terminate(true);
            MDVT_REQUIRE((int)rc.out.size() <= mdvt_ffv1::kHeaderStride, "slice header of %d bytes", (int)rc.out.size());
            memset(headers_host + si * mdvt_ffv1::kHeaderStride, 0, mdvt_ffv1::kHeaderStride);
            memcpy(headers_host + si * mdvt_ffv1::kHeaderStride, rc.out.data(), rc.out.size());
            header_len_host[si] = (int32_t)rc.out.size();
        }
    return MDVT_OK;
}

extern "C" int mdvt_ffv1_encode_frames(const uint8_t *frames, int64_t frame_stride, int64_t row_pitch, int n_frames, int width,
                                       int height, int nh, int nv, int alpha, int context_model, int bgr_order, const uint8_t *headers,
                                       const int32_t *header_len, void *states, uint8_t *slices, int64_t capacity, int32_t *sizes,
                                       int64_t *offsets, uint8_t *packed, void *stream) {
    if (int rc = mdvt::check_stream(width, height, nh, nv, alpha)) return rc;
    if (int rc = mdvt::check_model(context_model)) return rc;
    MDVT_REQUIRE(n_frames >= 0, "negative frame count");
    if (n_frames == 0) return MDVT_OK;
    MDVT_REQUIRE(capacity >= mdvt_ffv1_slice_capacity(width, height, nh, nv, alpha) && capacity % 16 == 0,
                 "slice capacity %lld is below mdvt_ffv1_slice_capacity or not a multiple of 16", (long long)capacity);
    MDVT_REQUIRE(row_pitch >= 3 * (int64_t)width && frame_stride >= row_pitch * height, "bad pitches");
    MDVT_REQUIRE(frames && headers && header_len && states && slices && sizes && offsets && packed, "NULL buffer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int per_frame = nh * nv;
    const int64_t total = (int64_t)n_frames * per_frame;
    MDVT_REQUIRE(total < (1LL << 30), "too many slices in one call");
    mdvt::ffv1_crc_table_kernel<<<1, 256, 0, s>>>();
    const int threads = 64;
    if (context_model == 2 && !alpha) {
        // The coder reads its pixels byte by byte through the L1 (60 M sector look-ups per 3840x1080 frame), so the L1 matters as
        // much as the resident warps: with the default carve-out 10 CTAs per SM leave it 60 KB (6.5 k frames/s), 12 CTAs 23 KB
        // (2.8 k); asking for half of the array (132 KB of shared memory = 6 CTAs, 121 KB of L1) gives 7.9-8.3 k
        // (profiles/r02_ffv1_state_experiments.txt).  MDVT_FFV1_CARVEOUT=<percent> overrides (tuning aid).
        static const int carve = getenv("MDVT_FFV1_CARVEOUT") ? atoi(getenv("MDVT_FFV1_CARVEOUT")) : 52;
        MDVT_CUDA_TRY(cudaFuncSetAttribute(mdvt::ffv1_encode_tiny_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    }
    if (context_model == 2 && !alpha)
        mdvt::ffv1_encode_tiny_kernel<64><<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(
            frames, frame_stride, row_pitch, n_frames, width, height, nh, nv, bgr_order ? 0 : 2, bgr_order ? 2 : 0, headers, header_len, slices,
            capacity, sizes);
    else
        mdvt::ffv1_encode_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(
            frames, frame_stride, row_pitch, n_frames, width, height, nh, nv, 3 + alpha, bgr_order ? 0 : 2, bgr_order ? 2 : 0, context_model,
            headers, header_len, static_cast<mdvt_ffv1::VlcState *>(states), slices, capacity, sizes);
    mdvt::ffv1_offsets_kernel<<<1, 1024, 0, s>>>(sizes, n_frames, per_frame, offsets);
    mdvt::ffv1_pack_kernel<<<(unsigned)total, 128, 0, s>>>(slices, capacity, sizes, offsets, packed);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}

extern "C" int mdvt_ffv1_parse_config(const uint8_t *config_host, int config_len, int width, int height, int *nh, int *nv, int *alpha,
                                      int *context_model) {
    MDVT_REQUIRE(config_host && nh && nv && alpha && context_model && config_len > 0, "NULL / empty configuration record");
    mdvt::RangeReader rr(config_host, config_len);
    uint8_t st[32];
    memset(st, 128, sizeof st);
    const int version = rr.get_symbol(st), micro = rr.get_symbol(st), coder = rr.get_symbol(st), colourspace = rr.get_symbol(st);
    const int bits = rr.get_symbol(st), chroma = rr.get_rac(st), hshift = rr.get_symbol(st), vshift = rr.get_symbol(st);
    const int transparency = rr.get_rac(st);
    const int h_slices = 1 + rr.get_symbol(st), v_slices = 1 + rr.get_symbol(st);
    const int table_sets = rr.get_symbol(st);           // quant table sets: 1 = one of this library's small models, 2 = libavcodec's
    (void)micro, (void)bits, (void)chroma, (void)hshift, (void)vshift;
    if (rr.bad || version != 3 || coder != 0 || colourspace != 1) {
        mdvt::set_error("not an FFV1 version 3 Golomb-Rice RGB stream (version %d, coder %d, colourspace %d)", version, coder, colourspace);
        return MDVT_ERR_UNSUPPORTED;
    }
    if (mdvt::check_stream(width, height, h_slices, v_slices, transparency) != MDVT_OK) return MDVT_ERR_UNSUPPORTED;
    // everything else (8 bit, quant tables, CRC, ...) must be exactly what this library writes for these parameters
    uint8_t own[64];
    int own_len = 0;
    std::vector<uint8_t> headers((size_t)h_slices * v_slices * mdvt_ffv1::kHeaderStride);
    std::vector<int32_t> lens((size_t)h_slices * v_slices);
    int model = table_sets == 1 ? 1 : 0;
    for (;;) {
        if (int rc = mdvt_ffv1_stream_setup(width, height, h_slices, v_slices, transparency, model, own, 64, &own_len, headers.data(),
                                            lens.data()))
            return rc;
        if ((own_len == config_len && memcmp(own, config_host, (size_t)own_len) == 0) || model != 1) break;
        model = 2;   // one table set that is not model 1's: the record must then be model 2's
    }
    if (own_len != config_len || memcmp(own, config_host, (size_t)own_len) != 0) {
        mdvt::set_error("FFV1 stream parameters differ from the ones this library writes (8 bit, its quant tables, CRC)");
        return MDVT_ERR_UNSUPPORTED;
    }
    *nh = h_slices;
    *nv = v_slices;
    *alpha = transparency;
    *context_model = model;
    return MDVT_OK;
}

extern "C" int mdvt_ffv1_decode_frames(const uint8_t *packets, const int64_t *packet_offsets, int n_frames, int width, int height, int nh,
                                       int nv, int alpha, int context_model, int bgr_order, const uint8_t *headers, const int32_t *header_len, void *states,
                                       int64_t *slice_offsets, uint8_t *frames, int64_t frame_stride, int64_t row_pitch, int32_t *status,
                                       void *stream) {
    if (int rc = mdvt::check_stream(width, height, nh, nv, alpha)) return rc;
    if (int rc = mdvt::check_model(context_model)) return rc;
    MDVT_REQUIRE(n_frames >= 0, "negative frame count");
    if (n_frames == 0) return MDVT_OK;
    MDVT_REQUIRE(row_pitch >= 3 * (int64_t)width && frame_stride >= row_pitch * height, "bad pitches");
    MDVT_REQUIRE(packets && packet_offsets && headers && header_len && states && slice_offsets && frames && status, "NULL buffer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int per_frame = nh * nv;
    const int64_t total = (int64_t)n_frames * per_frame;
    MDVT_REQUIRE(total < (1LL << 30), "too many slices in one call");
    mdvt::ffv1_crc_table_kernel<<<1, 256, 0, s>>>();
    mdvt::ffv1_index_kernel<<<(n_frames + 31) / 32, 32, 0, s>>>(packets, packet_offsets, n_frames, per_frame, slice_offsets, status);
    const int threads = 64;
    if (context_model == 2 && !alpha)
        mdvt::ffv1_decode_tiny_kernel<64><<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(
            packets, packet_offsets, slice_offsets, n_frames, width, height, nh, nv, bgr_order ? 0 : 2, bgr_order ? 2 : 0, headers, header_len, frames,
            frame_stride, row_pitch, status);
    else
        mdvt::ffv1_decode_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(
            packets, packet_offsets, slice_offsets, n_frames, width, height, nh, nv, 3 + alpha, bgr_order ? 0 : 2, bgr_order ? 2 : 0, context_model,
            headers, header_len, static_cast<mdvt_ffv1::VlcState *>(states), frames, frame_stride, row_pitch, status);
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}
