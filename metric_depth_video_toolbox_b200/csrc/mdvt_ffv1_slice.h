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

// Context models (the quant tables are part of the configuration record, any FFV1 decoder follows them):
//   0  libavcodec's set 0 for <= 8-bit content: quant11 on three differences, (11*11*11 + 1) / 2 = 666 contexts per plane
//      context -- what cv2.VideoWriter's files use; packets are byte-identical to libavcodec's at equal slice layout;
//   1  a 5-level table on the same three differences, (5*5*5 + 1) / 2 = 63 contexts: 1 KB of coder state per slice instead
//      of 10.6 KB, and less context dilution for slices of a few thousand samples;
//   2  a 3-level table (the sign of each difference), (3*3*3 + 1) / 2 = 14 contexts: 224 bytes of coder state per slice, which
//      the device encoder keeps in shared memory; ~1 % larger files than model 1 on film-like content.
constexpr int kContexts = 666;        // per plane context, model 0
constexpr int kContextsSmall = 63;
constexpr int kContextsTiny = 14;
MDVT_FFV1_HD inline int contexts_of(int model) { return model == 2 ? kContextsTiny : (model ? kContextsSmall : kContexts); }
constexpr int kHeaderStride = 16;     // bytes reserved per precomputed range-coded slice header
constexpr int kFooterBytes = 8;       // 3-byte size, error-status byte, CRC-32

// VLC state of one context (ffv1.h VlcState), 8 bytes, handled as one 64-bit word: one load + one store per coded sample.
// bits 0-31 error_sum, 32-47 drift (int16), 48-55 bias (int8), 56-63 count.
typedef uint64_t VlcState;
constexpr uint64_t kVlcInit = 4ull | (1ull << 56);   // error_sum 4, drift 0, bias 0, count 1

MDVT_FFV1_HD inline int clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}

// Big-endian bit writer into the slice's own (4-byte aligned) byte range, carrying the running CRC-32 (polynomial
// 0x04C11DB7, MSB first, initial value 0, no final xor -- libavutil's AV_CRC_32_IEEE as ffv1enc.c uses it) of everything
// emitted.  Whole 32-bit words leave the accumulator; the CRC takes a word per step through four tables
// (crc_table[k * 256 + i] = CRC of byte i followed by k zero bytes).
struct BitSink {
    uint8_t *out;
    uint32_t pos;       // bytes written
    uint32_t crc;
    uint64_t acc;       // pending bits, right-aligned
    int nbits;          // < 32 between calls
    const uint32_t *crc_table;

    MDVT_FFV1_HD inline void byte(uint32_t b) {
        out[pos++] = (uint8_t)b;
        crc = (crc << 8) ^ crc_table[(crc >> 24) ^ b];
    }
    MDVT_FFV1_HD inline void word(uint32_t v) {
#ifdef __CUDA_ARCH__
        *reinterpret_cast<uint32_t *>(out + pos) = __byte_perm(v, 0, 0x0123);
#else
        out[pos] = (uint8_t)(v >> 24);
        out[pos + 1] = (uint8_t)(v >> 16);
        out[pos + 2] = (uint8_t)(v >> 8);
        out[pos + 3] = (uint8_t)v;
#endif
        pos += 4;
        const uint32_t x = crc ^ v;
        crc = crc_table[768 + (x >> 24)] ^ crc_table[512 + ((x >> 16) & 0xFFu)] ^ crc_table[256 + ((x >> 8) & 0xFFu)] ^ crc_table[x & 0xFFu];
    }
    MDVT_FFV1_HD inline void put(int n, uint32_t v) {   // n <= 32, v < 2^n
        acc = (acc << n) | v;
        nbits += n;
        if (nbits >= 32) {
            nbits -= 32;
            word((uint32_t)(acc >> nbits));
        }
    }
    MDVT_FFV1_HD inline void flush() {   // pads the last byte with zero bits
        while (nbits >= 8) {
            nbits -= 8;
            byte((uint32_t)(acc >> nbits) & 0xFFu);
        }
        if (nbits) {
            byte((uint32_t)(acc << (8 - nbits)) & 0xFFu);
            nbits = 0;
        }
    }
};

// The quantiser of a neighbour difference as arithmetic.  Model 0: libavcodec's quant11[] (levels change at 1, 2, 5, 12,
// 35); model 1: 5 levels (changes at 1 and 4, the shape of libavcodec's quant5[]); model 2: 3 levels (the sign).
template <int SMALL>
MDVT_FFV1_HD inline int quant(int d) {
    d &= 0xFF;
    const int neg = d >= 128;
    const int m = neg ? 256 - d : d;
    const int q = SMALL == 2 ? (m >= 1) : (SMALL ? (m >= 1) + (m >= 4) : (m >= 1) + (m >= 2) + (m >= 5) + (m >= 12) + (m >= 35));
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

// One sample through the adaptive Golomb-Rice coder (ffv1enc.c put_vlc_symbol + golomb.h set_ur_golomb, limit 12, escape
// 9); returns the updated state.
MDVT_FFV1_HD inline VlcState put_vlc(BitSink &bs, VlcState st, int v) {
    const uint32_t error_sum = (uint32_t)st;
    const uint32_t hi = (uint32_t)(st >> 32);
    int drift = (int16_t)(hi & 0xFFFFu), bias = (int8_t)((hi >> 16) & 0xFFu), count = (int)(hi >> 24);
    v = fold9(v - bias);
    // k = smallest k with (count << k) >= error_sum  (the `while (i < error_sum) { k++; i += i; }` of put_vlc_symbol)
    int k = clz32((uint32_t)count) - clz32(error_sum);
    if (k < 0) k = 0;
    else if (((uint32_t)count << k) < error_sum) ++k;
    const int code = (2 * drift + count) < 0 ? ~v : v;
    const uint32_t u = code >= 0 ? 2u * (uint32_t)code : (uint32_t)(-2 * code - 1);
    const uint32_t e = u >> k;
    if (e < 12)
        bs.put((int)e + k + 1, (1u << k) + (u & ((1u << k) - 1u)));
    else
        bs.put(21, u - 11u);
    drift += v;
    uint32_t es = error_sum + (uint32_t)(v < 0 ? -v : v);
    if (count == 128) {
        count >>= 1;
        drift >>= 1;
        es >>= 1;
    }
    ++count;
    if (drift <= -count) {
        bias = bias > -128 ? bias - 1 : -128;
        drift += count;
        if (drift < -count + 1) drift = -count + 1;
    } else if (drift > 0) {
        bias = bias < 127 ? bias + 1 : 127;
        drift -= count;
        if (drift > 0) drift = 0;
    }
    return (uint64_t)es | ((uint64_t)((uint32_t)(drift & 0xFFFF) | ((uint32_t)(bias & 0xFF) << 16) | ((uint32_t)count << 24)) << 32);
}

// The bytes of one source pixel a plane needs, and the plane's sample from them.  Planes: 0: G' = G + ((B-G + R-G) >> 2),
// 1: B - G + 256, 2: R - G + 256, 3: alpha = 255.  ib / ir: byte positions of blue and red (2 / 0 for RGB-order frames,
// 0 / 2 for BGR).  Loading and evaluating are separate so that the loads can be issued an iteration ahead of their use.
struct Raw {
    int g, b, r;
};

template <int PL>
MDVT_FFV1_HD inline Raw load_raw(const uint8_t *px, int ib, int ir) {
    Raw v;
    v.g = v.b = v.r = 0;
    if (PL < 3) v.g = MDVT_FFV1_LD(px + 1);
    if (PL == 0 || PL == 1) v.b = MDVT_FFV1_LD(px + ib);
    if (PL == 0 || PL == 2) v.r = MDVT_FFV1_LD(px + ir);
    return v;
}

template <int PL>
MDVT_FFV1_HD inline int eval_raw(const Raw &v) {
    if (PL == 3) return 255;
    if (PL == 1) return v.b - v.g + 256;
    if (PL == 2) return v.r - v.g + 256;
    return v.g + ((v.b + v.r - 2 * v.g) >> 2);
}

// Context and prediction residual of one sample from its neighbours (ffv1.h get_context / predict, sign folded as in
// encode_line).  q_lt_t = quant(LT - T) comes from the previous sample (it was its quant(T - RT)).
template <int SMALL>
MDVT_FFV1_HD inline void context_and_residual(int L, int LT, int T, int RT, int C, int q_lt_t, int &q_t_rt, int &ctx, int &diff) {
    q_t_rt = quant<SMALL>(T - RT);
    ctx = quant<SMALL>(L - LT) + (SMALL == 2 ? 3 : (SMALL ? 5 : 11)) * q_lt_t + (SMALL == 2 ? 9 : (SMALL ? 25 : 121)) * q_t_rt;
    diff = C - median3(L, L + T - LT, T);
    if (ctx < 0) {
        ctx = -ctx;
        diff = -diff;
    }
    diff = fold9(diff);
}

struct SliceJob {
    const uint8_t *frame;     // first byte of the slice's top-left pixel
    int64_t row_pitch;        // bytes between rows of the frame
    int w, h;                 // slice size in pixels
    int n_planes;             // 3, or 4 with the constant alpha plane OpenCV's BGRA input produces
    int ib, ir;
    const uint8_t *header;    // range-coded slice header (mdvt_ffv1_stream_setup), header_len bytes
    int header_len;
    VlcState *states;         // n_plane_contexts * contexts_of(model), reset here (every frame is a key frame)
    uint8_t *out;             // 4-byte aligned, capacity >= slice_capacity(w, h, n_planes)
    const uint32_t *crc_table;   // 4 x 256 entries (BitSink)
};

// One line of one plane (ffv1enc.c encode_line).  The loop is software-pipelined by hand: while sample x is coded, the
// context / residual of sample x+1 are already known and its VLC state is in flight, and the pixel bytes of sample x+2
// are in flight -- the serial coder never waits for memory except when two consecutive samples share a context, and
// then the state is forwarded in registers.  first1 / first2: sample 0 of rows y-1 / y-2 (the format's left border:
// ffv1enc.c encode_rgb_frame sets sample[p][0][-1] = sample[p][1][0] and sample[p][1][w] = sample[p][1][w-1]; rows above
// the slice read as 0).
// STRIDE: distance, in states, between consecutive contexts of this slice (1: a contiguous block per slice; > 1: the states of
// several slices interleaved, e.g. in shared memory with one slice per thread).
template <int PL, int SMALL, int STRIDE = 1>
MDVT_FFV1_HD inline void encode_line(BitSink &bs, VlcState *states, const uint8_t *row, const uint8_t *up, bool has_up, int w, int ib,
                                     int ir, int &run_index, int &first1, int &first2) {
    const uint8_t *upper = has_up ? up : row;   // a readable address either way; the value is dropped without a row above
    const int last = w - 1;
    int C = eval_raw<PL>(load_raw<PL>(row, ib, ir));
    int T = first1;
    int RT = has_up ? eval_raw<PL>(load_raw<PL>(upper + 3 * (last < 1 ? last : 1), ib, ir)) : 0;
    int q, ctx, diff;
    context_and_residual<SMALL>(first1, first2, T, RT, C, quant<SMALL>(first2 - T), q, ctx, diff);
    first2 = first1;
    first1 = C;
    VlcState *sp = states + ctx * STRIDE;
    VlcState st = *sp;
    // bytes of sample 1 (current row x = 1, upper row x = 2)
    Raw rc_n = load_raw<PL>(row + 3 * (last < 1 ? last : 1), ib, ir);
    Raw rrt_n = load_raw<PL>(upper + 3 * (last < 2 ? last : 2), ib, ir);
    int run_count = 0, run_mode = 0;
    for (int x = 0; x < w; ++x) {
        // bytes of sample x + 2: issued now, used in the next iteration
        const int xc = x + 2 < last ? x + 2 : last, xu = x + 3 < last ? x + 3 : last;
        const Raw rc_nn = load_raw<PL>(row + 3 * xc, ib, ir);
        const Raw rrt_nn = load_raw<PL>(upper + 3 * xu, ib, ir);
        // sample x + 1: context, residual, state load
        int ctx_n = ctx, diff_n = 0;
        VlcState *sp_n = sp;
        VlcState st_n = st;
        if (x < last) {
            const int Cn = eval_raw<PL>(rc_n);
            const int RTn = has_up ? eval_raw<PL>(rrt_n) : 0;
            int qn;
            context_and_residual<SMALL>(C, T, RT, RTn, Cn, q, qn, ctx_n, diff_n);
            q = qn;
            T = RT;
            RT = RTn;
            C = Cn;
            sp_n = states + ctx_n * STRIDE;
            st_n = *sp_n;
        }
        // sample x
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
        if (!run_mode) {
            st = put_vlc(bs, st, diff);
            *sp = st;
            if (sp_n == sp) st_n = st;   // the load above was issued before this store
        }
        sp = sp_n;
        st = st_n;
        ctx = ctx_n;
        diff = diff_n;
        rc_n = rc_nn;
        rrt_n = rrt_nn;
    }
    if (run_mode) {
        while (run_count >= (1 << log2_run(run_index))) {
            run_count -= 1 << log2_run(run_index);
            ++run_index;
            bs.put(1, 1);
        }
        if (run_count) bs.put(1, 1);
    }
}

// Codes one slice; returns its size in the packet (body + footer).
template <int SMALL, int STRIDE = 1>
MDVT_FFV1_HD inline uint32_t encode_slice(const SliceJob &job) {
    constexpr int NC = SMALL == 2 ? kContextsTiny : (SMALL ? kContextsSmall : kContexts);
    BitSink bs;
    bs.out = job.out;
    bs.pos = 0;
    bs.crc = 0;
    bs.acc = 0;
    bs.nbits = 0;
    bs.crc_table = job.crc_table;
    for (int i = 0; i < job.header_len; ++i) bs.put(8, job.header[i]);

    const int n_pc = job.n_planes > 3 ? 3 : 2;   // plane contexts: G | B,R | alpha
    for (int i = 0; i < n_pc * NC; ++i) job.states[i * STRIDE] = kVlcInit;

    int run_index = 0;
    int f1_0 = 0, f1_1 = 0, f1_2 = 0, f1_3 = 0, f2_0 = 0, f2_1 = 0, f2_2 = 0, f2_3 = 0;   // sample 0 of rows y-1, y-2 per plane
    for (int y = 0; y < job.h; ++y) {
        const uint8_t *row = job.frame + (int64_t)y * job.row_pitch;
        const uint8_t *up = row - job.row_pitch;
        const bool has_up = y > 0;
        encode_line<0, SMALL, STRIDE>(bs, job.states, row, up, has_up, job.w, job.ib, job.ir, run_index, f1_0, f2_0);
        encode_line<1, SMALL, STRIDE>(bs, job.states + NC * STRIDE, row, up, has_up, job.w, job.ib, job.ir, run_index, f1_1, f2_1);
        encode_line<2, SMALL, STRIDE>(bs, job.states + NC * STRIDE, row, up, has_up, job.w, job.ib, job.ir, run_index, f1_2, f2_2);
        if (job.n_planes > 3) encode_line<3, SMALL, STRIDE>(bs, job.states + 2 * NC * STRIDE, row, up, has_up, job.w, job.ib, job.ir, run_index, f1_3, f2_3);
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

// ---- decoder: the mirror image (ffv1dec.c decode_slice -> decode_rgb_frame -> decode_line / get_vlc_symbol) -------------------
// Reads the slices this encoder writes (every frame a key frame, slice header known in advance, quant-table set 0).

// Big-endian bit reader over [p, end): a 64-bit window refilled 32 bits at a time; reads past the end see zero bits.
struct BitSource {
    const uint8_t *p, *end;
    uint64_t acc;   // valid bits left-aligned
    int nbits;

    MDVT_FFV1_HD inline void refill() {   // at least 33 valid bits afterwards
        if (nbits <= 32) {
            uint32_t w = 0;
            if (p + 4 <= end) {
                w = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
            } else {
                for (int i = 0; i < 4; ++i)
                    if (p + i < end) w |= (uint32_t)p[i] << (24 - 8 * i);
            }
            p += 4;
            acc |= (uint64_t)w << (32 - nbits);
            nbits += 32;
        }
    }
    MDVT_FFV1_HD inline uint32_t peek32() {
        refill();
        return (uint32_t)(acc >> 32);
    }
    MDVT_FFV1_HD inline void skip(int n) {   // n <= 32 after a peek32 / refill
        acc <<= n;
        nbits -= n;
    }
    MDVT_FFV1_HD inline uint32_t get(int n) {   // 0 <= n <= 25
        if (n == 0) return 0;
        const uint32_t v = peek32() >> (32 - n);
        skip(n);
        return v;
    }
};

// get_vlc_symbol (ffv1dec.c) + get_ur_golomb (golomb.h, limit 12, escape 9): returns the residual, updates the state.
MDVT_FFV1_HD inline int get_vlc(BitSource &bs, VlcState &st) {
    const uint32_t error_sum = (uint32_t)st;
    const uint32_t hi = (uint32_t)(st >> 32);
    int drift = (int16_t)(hi & 0xFFFFu), bias = (int8_t)((hi >> 16) & 0xFFu), count = (int)(hi >> 24);
    int k = clz32((uint32_t)count) - clz32(error_sum);
    if (k < 0) k = 0;
    else if (((uint32_t)count << k) < error_sum) ++k;
    const uint32_t window = bs.peek32();
    const int z = clz32(window);          // leading zero bits (32 if the window is empty)
    uint32_t u;
    if (z < 12) {
        bs.skip(z + 1);
        u = ((uint32_t)z << k) + bs.get(k);
    } else {
        bs.skip(12);
        u = bs.get(9) + 11u;
    }
    int v = (int)(u >> 1) ^ -(int)(u & 1u);
    if ((2 * drift + count) < 0) v = ~v;
    const int ret = fold9(v + bias);
    drift += v;
    uint32_t es = error_sum + (uint32_t)(v < 0 ? -v : v);
    if (count == 128) {
        count >>= 1;
        drift >>= 1;
        es >>= 1;
    }
    ++count;
    if (drift <= -count) {
        bias = bias > -128 ? bias - 1 : -128;
        drift += count;
        if (drift < -count + 1) drift = -count + 1;
    } else if (drift > 0) {
        bias = bias < 127 ? bias + 1 : 127;
        drift -= count;
        if (drift > 0) drift = 0;
    }
    st = (uint64_t)es | ((uint64_t)((uint32_t)(drift & 0xFFFF) | ((uint32_t)(bias & 0xFF) << 16) | ((uint32_t)count << 24)) << 32);
    return ret;
}

// Plain (coherent) loads: the frame being decoded is written by the same thread.
template <int PL>
MDVT_FFV1_HD inline int plane_of_pixel(const uint8_t *px, int ib, int ir) {
    Raw v;
    v.g = v.b = v.r = 0;
    if (PL < 3) v.g = px[1];
    if (PL == 0 || PL == 1) v.b = px[ib];
    if (PL == 0 || PL == 2) v.r = px[ir];
    return eval_raw<PL>(v);
}

struct SliceInput {
    const uint8_t *data;      // the slice's bytes in the packet: header, Golomb-Rice bits, footer
    uint32_t size;            // including the 8 footer bytes
    const uint8_t *header;    // what the header must be (mdvt_ffv1_stream_setup)
    int header_len;
    uint8_t *frame;           // first byte of the slice's top-left pixel in the output frame (u8x3)
    int64_t row_pitch;
    int w, h, n_planes, ib, ir;
    VlcState *states;
    const uint32_t *crc_table;   // 256 entries (the first of BitSink's four tables)
};

// One line of one plane.  Plane 0 / 1 leave their samples in the output row as scratch (G' in the green byte; the low 8
// bits of B' in the blue byte and its ninth bit in the red byte); plane 2 turns the three into the final pixel (inverse
// RCT); plane 3 (alpha) is decoded and dropped.  Neighbours of the row above are recomputed from its final pixels.
template <int PL, int SMALL, int STRIDE = 1>
MDVT_FFV1_HD inline void decode_line(BitSource &bs, VlcState *states, uint8_t *row, const uint8_t *up, bool has_up, int w, int ib, int ir,
                                     int &run_index, int &first1, int &first2) {
    const int last = w - 1;
    int L = first1, LT = first2, T = first1;
    int RT = has_up ? plane_of_pixel<PL>(up + 3 * (last < 1 ? last : 1), ib, ir) : 0;
    int q_lt_t = quant<SMALL>(LT - T);
    int run_count = 0, run_mode = 0;
    int row_first = 0;
    for (int x = 0; x < w; ++x) {
        // the next sample's upper-right neighbour: a load that does not depend on the serial chain, issued a whole sample early
        const int xr = x + 2 < last ? x + 2 : last;
        const int RT_next = has_up ? plane_of_pixel<PL>(up + 3 * xr, ib, ir) : 0;
        const int q_t_rt = quant<SMALL>(T - RT);
        int ctx = quant<SMALL>(L - LT) + (SMALL == 2 ? 3 : (SMALL ? 5 : 11)) * q_lt_t + (SMALL == 2 ? 9 : (SMALL ? 25 : 121)) * q_t_rt;
        const bool sign = ctx < 0;
        if (sign) ctx = -ctx;
        int diff;
        if (ctx == 0 && run_mode == 0) run_mode = 1;
        if (run_mode) {
            if (run_count == 0 && run_mode == 1) {
                if (bs.get(1)) {
                    run_count = 1 << log2_run(run_index);
                    if (x + run_count <= w) ++run_index;
                } else {
                    run_count = (int)bs.get(log2_run(run_index));
                    if (run_index) --run_index;
                    run_mode = 2;
                }
            }
            --run_count;
            if (run_count < 0) {
                run_mode = 0;
                run_count = 0;
                diff = get_vlc(bs, states[ctx * STRIDE]);
                if (diff >= 0) ++diff;
            } else {
                diff = 0;
            }
        } else {
            diff = get_vlc(bs, states[ctx * STRIDE]);
        }
        if (sign) diff = -diff;
        const int cur = (median3(L, L + T - LT, T) + diff) & 511;
        if (x == 0) row_first = cur;
        uint8_t *px = row + 3 * x;
        if (PL == 0) {
            px[1] = (uint8_t)cur;
        } else if (PL == 1) {
            px[ib] = (uint8_t)cur;
            px[ir] = (uint8_t)(cur >> 8);
        } else if (PL == 2) {
            const int b = ((int)px[ib] | ((int)px[ir] << 8)) - 256, r = cur - 256;
            const int g = (int)px[1] - ((b + r) >> 2);
            px[1] = (uint8_t)g;
            px[ib] = (uint8_t)(b + g);
            px[ir] = (uint8_t)(r + g);
        }
        LT = T;
        T = RT;
        L = cur;
        q_lt_t = q_t_rt;
        RT = RT_next;
    }
    first2 = first1;
    first1 = row_first;
}

// Decodes one slice into the frame; returns 0, or a negative code: -1 the header is not the expected one, -2 the size in
// the footer does not match, -3 the bit stream ran past the slice, -4 the slice's CRC-32 is wrong (checked first: a damaged
// slice is reported, not decoded -- ffv1dec.c decode_frame does the same check on every slice when ec is set).
template <int SMALL, int STRIDE = 1>
MDVT_FFV1_HD inline int decode_slice(const SliceInput &in) {
    constexpr int NC = SMALL == 2 ? kContextsTiny : (SMALL ? kContextsSmall : kContexts);
    if (in.size < (uint32_t)(in.header_len + kFooterBytes)) return -2;
    for (int i = 0; i < in.header_len; ++i)
        if (in.data[i] != in.header[i]) return -1;
    const uint8_t *foot = in.data + in.size - kFooterBytes;
    const uint32_t body = ((uint32_t)foot[0] << 16) | ((uint32_t)foot[1] << 8) | (uint32_t)foot[2];
    if (body + kFooterBytes != in.size) return -2;
    uint32_t crc = 0;   // over body, size, error status and the stored CRC: zero for an intact slice
    for (uint32_t i = 0; i < in.size; ++i) crc = (crc << 8) ^ in.crc_table[(crc >> 24) ^ in.data[i]];
    if (crc != 0) return -4;
    BitSource bs;
    bs.p = in.data + in.header_len;
    bs.end = foot;
    bs.acc = 0;
    bs.nbits = 0;
    const int n_pc = in.n_planes > 3 ? 3 : 2;
    for (int i = 0; i < n_pc * NC; ++i) in.states[i * STRIDE] = kVlcInit;
    int run_index = 0;
    int f1_0 = 0, f1_1 = 0, f1_2 = 0, f1_3 = 0, f2_0 = 0, f2_1 = 0, f2_2 = 0, f2_3 = 0;
    for (int y = 0; y < in.h; ++y) {
        uint8_t *row = in.frame + (int64_t)y * in.row_pitch;
        const uint8_t *up = row - in.row_pitch;
        const bool has_up = y > 0;
        decode_line<0, SMALL, STRIDE>(bs, in.states, row, up, has_up, in.w, in.ib, in.ir, run_index, f1_0, f2_0);
        decode_line<1, SMALL, STRIDE>(bs, in.states + NC * STRIDE, row, up, has_up, in.w, in.ib, in.ir, run_index, f1_1, f2_1);
        decode_line<2, SMALL, STRIDE>(bs, in.states + NC * STRIDE, row, up, has_up, in.w, in.ib, in.ir, run_index, f1_2, f2_2);
        if (in.n_planes > 3) decode_line<3, SMALL, STRIDE>(bs, in.states + 2 * NC * STRIDE, row, up, has_up, in.w, in.ib, in.ir, run_index, f1_3, f2_3);
    }
    // bits consumed must lie inside the body (the window holds bytes read ahead)
    const int64_t consumed_bits = (int64_t)(bs.p - (in.data + in.header_len)) * 8 - bs.nbits;
    if (consumed_bits > (int64_t)(foot - (in.data + in.header_len)) * 8) return -3;
    return 0;
}

// Worst case of one slice: 21 bits per sample (escape code) + 1 run bit, 512 bits for the run-length prefixes of a
// descending run index, header, footer, padding to 16 bytes.
MDVT_FFV1_HD inline int64_t slice_capacity(int w, int h, int n_planes) {
    const int64_t bits = (int64_t)w * h * n_planes * 22 + 512;
    return ((bits + 7) / 8 + kHeaderStride + kFooterBytes + 15) / 16 * 16;
}

}  // namespace mdvt_ffv1
