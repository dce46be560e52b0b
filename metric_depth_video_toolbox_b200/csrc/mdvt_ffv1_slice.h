// FFV1 version 3 slice coder (Golomb-Rice, RGB with the JPEG2000 RCT, 9-bit samples, optional constant alpha plane):
// the sequential per-slice algorithm, written once and compiled as device code by mdvt_ffv1.cu (one thread per slice).
// The same text compiles as plain C++ (MDVT_FFV1_HD empty) -- tests/ build it that way to step it against
// oracle/ffv1_oracle.py on the CPU; nothing in the product calls a host build of it.
//
// What it replaces: the entropy coding behind every `cv2.VideoWriter(..., fourcc "FFV1", ...)` of the reference
// (stereo_rerender.py:420-442,941, depth_frames_helper.py:125-161, 3d_view_depthfile.py:118-127), i.e. libavcodec's
// ffv1enc.c encode_slice -> encode_rgb_frame -> encode_line / put_vlc_symbol (RFC 9043 sections 3.3-3.8, 4.5-4.7).
#pragma once

#include <stdint.h>

#ifndef MDVT_FFV1_HD
#define MDVT_FFV1_HD
#endif
#ifndef MDVT_FFV1_LD
#define MDVT_FFV1_LD(p) (*(p))
#endif

namespace mdvt_ffv1 {

constexpr int kContexts = 666;        // (11*11*11 + 1) / 2: quant-table set 0 of libavcodec for <= 8-bit content
constexpr int kHeaderStride = 16;     // bytes reserved per precomputed range-coded slice header
constexpr int kFooterBytes = 8;       // 3-byte size, error-status byte, CRC-32

// VLC state of one context (ffv1.h VlcState): 8 bytes, one load + one store per coded sample.
struct VlcState {
    uint32_t error_sum;
    int16_t drift;
    int8_t bias;
    uint8_t count;
};

// Big-endian bit writer into the slice's own byte range, carrying the running CRC-32 (polynomial 0x04C11DB7, MSB first,
// initial value 0, no final xor -- libavutil's AV_CRC_32_IEEE as ffv1enc.c uses it) of every byte it has emitted.
struct BitSink {
    uint8_t *out;
    uint32_t pos;       // bytes written
    uint32_t crc;
    uint64_t acc;       // pending bits, right-aligned
    int nbits;
    const uint32_t *crc_table;

    MDVT_FFV1_HD inline void byte(uint32_t b) {
        out[pos++] = (uint8_t)b;
        crc = (crc << 8) ^ crc_table[(crc >> 24) ^ b];
    }
    MDVT_FFV1_HD inline void put(int n, uint32_t v) {   // n <= 32
        acc = (acc << n) | v;
        nbits += n;
        while (nbits >= 8) {
            nbits -= 8;
            byte((uint32_t)(acc >> nbits) & 0xFFu);
        }
    }
    MDVT_FFV1_HD inline void flush() {
        if (nbits) {
            byte((uint32_t)(acc << (8 - nbits)) & 0xFFu);
            nbits = 0;
        }
    }
};

MDVT_FFV1_HD inline int quant11(int d) {   // libavcodec's quant11[] as arithmetic: levels change at 1, 2, 5, 12, 35
    d &= 0xFF;
    const int neg = d >= 128;
    const int m = neg ? 256 - d : d;
    const int q = (m >= 1) + (m >= 2) + (m >= 5) + (m >= 12) + (m >= 35);
    return neg ? -q : q;
}

MDVT_FFV1_HD inline int fold9(int v) {
    v &= 511;
    return v >= 256 ? v - 512 : v;
}

MDVT_FFV1_HD inline int median3(int a, int b, int c) {
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return c < lo ? lo : (c > hi ? hi : c);
}

MDVT_FFV1_HD inline int log2_run(int i) {   // ffv1.h ff_log2_run[41]
    return i < 16 ? i >> 2 : (i < 24 ? 4 + ((i - 16) >> 1) : i - 16);
}

// One sample through the adaptive Golomb-Rice coder (ffv1enc.c put_vlc_symbol + golomb.h set_ur_golomb, limit 12, escape 9).
MDVT_FFV1_HD inline void put_vlc(BitSink &bs, VlcState *sp, int v) {
    VlcState s = *sp;
    v = fold9(v - s.bias);
    int i = s.count, k = 0;
    while (i < (int)s.error_sum) {
        ++k;
        i += i;
    }
    const int code = (2 * s.drift + s.count) < 0 ? ~v : v;
    const uint32_t u = code >= 0 ? 2u * (uint32_t)code : (uint32_t)(-2 * code - 1);
    const uint32_t e = u >> k;
    if (e < 12)
        bs.put((int)e + k + 1, (1u << k) + (u & ((1u << k) - 1u)));
    else
        bs.put(21, u - 11u);
    int drift = s.drift + v, count = s.count;
    uint32_t es = s.error_sum + (uint32_t)(v < 0 ? -v : v);
    if (count == 128) {
        count >>= 1;
        drift >>= 1;
        es >>= 1;
    }
    ++count;
    int bias = s.bias;
    if (drift <= -count) {
        bias = bias > -128 ? bias - 1 : -128;
        drift += count;
        if (drift < -count + 1) drift = -count + 1;
    } else if (drift > 0) {
        bias = bias < 127 ? bias + 1 : 127;
        drift -= count;
        if (drift > 0) drift = 0;
    }
    s.error_sum = es;
    s.drift = (int16_t)drift;
    s.bias = (int8_t)bias;
    s.count = (uint8_t)count;
    *sp = s;
}

// Sample of plane `pl` (0: G', 1: B - G + 256, 2: R - G + 256, 3: alpha = 255) of one source pixel (three bytes; ib / ir
// are the byte positions of blue and red: 2 / 0 for RGB-order frames, 0 / 2 for BGR).
MDVT_FFV1_HD inline int plane_sample(const uint8_t *px, int pl, int ib, int ir) {
    if (pl == 3) return 255;
    const int g = MDVT_FFV1_LD(px + 1);
    if (pl == 1) return (int)MDVT_FFV1_LD(px + ib) - g + 256;
    if (pl == 2) return (int)MDVT_FFV1_LD(px + ir) - g + 256;
    const int b = (int)MDVT_FFV1_LD(px + ib) - g, r = (int)MDVT_FFV1_LD(px + ir) - g;
    return g + ((b + r) >> 2);
}

struct SliceJob {
    const uint8_t *frame;     // first byte of the slice's top-left pixel
    int64_t row_pitch;        // bytes between rows of the frame
    int w, h;                 // slice size in pixels
    int n_planes;             // 3, or 4 with the constant alpha plane OpenCV's BGRA input produces
    int ib, ir;
    const uint8_t *header;    // range-coded slice header (mdvt_ffv1_stream_setup), header_len bytes
    int header_len;
    VlcState *states;         // n_plane_contexts * kContexts, reset here (every frame is a key frame)
    uint8_t *out;             // capacity >= mdvt_ffv1_slice_capacity(w, h, n_planes)
    const uint32_t *crc_table;
};

// Codes one slice; returns its size in the packet (body + footer).
MDVT_FFV1_HD inline uint32_t encode_slice(const SliceJob &job) {
    BitSink bs;
    bs.out = job.out;
    bs.pos = 0;
    bs.crc = 0;
    bs.acc = 0;
    bs.nbits = 0;
    bs.crc_table = job.crc_table;
    for (int i = 0; i < job.header_len; ++i) bs.byte(job.header[i]);

    const int n_pc = job.n_planes > 3 ? 3 : 2;   // plane contexts: G | B,R | alpha
    for (int i = 0; i < n_pc * kContexts; ++i) {
        VlcState z;
        z.error_sum = 4;
        z.drift = 0;
        z.bias = 0;
        z.count = 1;
        job.states[i] = z;
    }

    int run_index = 0;
    int first1[4] = {0, 0, 0, 0}, first2[4] = {0, 0, 0, 0};   // sample 0 of rows y-1 and y-2 per plane
    const int w = job.w;
    for (int y = 0; y < job.h; ++y) {
        const uint8_t *row = job.frame + (int64_t)y * job.row_pitch;
        const uint8_t *up = row - job.row_pitch;
        for (int pl = 0; pl < job.n_planes; ++pl) {
            VlcState *states = job.states + ((pl + 1) >> 1) * kContexts;
            // neighbours of x = 0 (ffv1enc.c encode_rgb_frame: sample[p][0][-1] = sample[p][1][0];
            // sample[p][1][w] = sample[p][1][w-1]; rows above the slice read as 0)
            int L = first1[pl], LT = first2[pl], T = first1[pl];
            int RT = y ? plane_sample(up + 3 * (w > 1 ? 1 : 0), pl, job.ib, job.ir) : 0;
            int run_count = 0, run_mode = 0;
            int row_first = 0;
            for (int x = 0; x < w; ++x) {
                const int cur = plane_sample(row + 3 * x, pl, job.ib, job.ir);
                if (x == 0) row_first = cur;
                int ctx = quant11(L - LT) + 11 * quant11(LT - T) + 121 * quant11(T - RT);
                int diff = cur - median3(L, L + T - LT, T);
                if (ctx < 0) {
                    ctx = -ctx;
                    diff = -diff;
                }
                diff = fold9(diff);
                if (ctx == 0) run_mode = 1;
                if (run_mode) {
                    if (diff) {
                        while (run_count >= (1 << log2_run(run_index))) {
                            run_count -= 1 << log2_run(run_index);
                            ++run_index;
                            bs.put(1, 1);
                        }
                        bs.put(1 + log2_run(run_index), (uint32_t)run_count);
                        if (run_index) --run_index;
                        run_count = 0;
                        run_mode = 0;
                        if (diff > 0) --diff;
                    } else {
                        ++run_count;
                    }
                }
                if (!run_mode) put_vlc(bs, states + ctx, diff);
                // slide the window
                LT = T;
                T = RT;
                L = cur;
                const int xr = x + 2 < w ? x + 2 : w - 1;
                RT = y ? plane_sample(up + 3 * xr, pl, job.ib, job.ir) : 0;
            }
            if (run_mode) {
                while (run_count >= (1 << log2_run(run_index))) {
                    run_count -= 1 << log2_run(run_index);
                    ++run_index;
                    bs.put(1, 1);
                }
                if (run_count) bs.put(1, 1);
            }
            first2[pl] = first1[pl];
            first1[pl] = row_first;
        }
    }
    bs.flush();
    const uint32_t body = bs.pos;
    bs.byte((body >> 16) & 0xFFu);
    bs.byte((body >> 8) & 0xFFu);
    bs.byte(body & 0xFFu);
    bs.byte(0);   // error status
    const uint32_t crc = bs.crc;
    job.out[bs.pos++] = (uint8_t)(crc >> 24);
    job.out[bs.pos++] = (uint8_t)(crc >> 16);
    job.out[bs.pos++] = (uint8_t)(crc >> 8);
    job.out[bs.pos++] = (uint8_t)crc;
    return bs.pos;
}

// Worst case of one slice: 21 bits per sample (escape code) + 1 run bit, 512 bits for the run-length prefixes of a
// descending run index, header, footer, padding to 16 bytes.
MDVT_FFV1_HD inline int64_t slice_capacity(int w, int h, int n_planes) {
    const int64_t bits = (int64_t)w * h * n_planes * 22 + 512;
    return ((bits + 7) / 8 + kHeaderStride + kFooterBytes + 15) / 16 * 16;
}

}  // namespace mdvt_ffv1
