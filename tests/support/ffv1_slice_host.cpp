// Test support only: csrc/mdvt_ffv1_slice.h (the per-slice FFV1 coder the device runs, one thread per slice) compiled
// as plain C++, so that tests can step the very same text against oracle/ffv1_oracle.py and libavcodec without a GPU.
// Nothing in the package builds or loads this.
#include <cstdlib>
#include <cstring>

#include "mdvt_ffv1_slice.h"

static uint32_t g_crc[1024];   // [k * 256 + i]: CRC of byte i followed by k zero bytes
static void init_crc() {
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i << 24;
        for (int k = 0; k < 8; ++k) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : c << 1;
        g_crc[i] = c;
    }
    for (int k = 1; k < 4; ++k)
        for (uint32_t i = 0; i < 256; ++i) g_crc[k * 256 + i] = (g_crc[(k - 1) * 256 + i] << 8) ^ g_crc[g_crc[(k - 1) * 256 + i] >> 24];
}

extern "C" long long ffv1_host_slice_capacity(int w, int h, int n_planes) { return mdvt_ffv1::slice_capacity(w, h, n_planes); }

// Encodes all nh x nv slices of one frame into `packet` (slices concatenated in raster order); returns its size.
extern "C" long long ffv1_host_encode_frame(const uint8_t *frame, long long row_pitch, int width, int height, int nh, int nv,
                                            int n_planes, int bgr_order, int context_model, const uint8_t *headers,
                                            const int32_t *header_len, uint8_t *packet, long long packet_capacity) {
    init_crc();
    mdvt_ffv1::VlcState *states = (mdvt_ffv1::VlcState *)malloc(sizeof(mdvt_ffv1::VlcState) * 3 * mdvt_ffv1::kContexts);
    long long pos = 0;
    for (int sy = 0; sy < nv; ++sy)
        for (int sx = 0; sx < nh; ++sx) {
            const int si = sy * nh + sx;
            const int x0 = (int)((long long)sx * width / nh), x1 = (int)((long long)(sx + 1) * width / nh);
            const int y0 = (int)((long long)sy * height / nv), y1 = (int)((long long)(sy + 1) * height / nv);
            mdvt_ffv1::SliceJob job;
            job.frame = frame + y0 * row_pitch + 3LL * x0;
            job.row_pitch = row_pitch;
            job.w = x1 - x0;
            job.h = y1 - y0;
            job.n_planes = n_planes;
            job.ib = bgr_order ? 0 : 2;
            job.ir = bgr_order ? 2 : 0;
            job.header = headers + si * mdvt_ffv1::kHeaderStride;
            job.header_len = header_len[si];
            job.states = states;
            if (pos + mdvt_ffv1::slice_capacity(job.w, job.h, n_planes) > packet_capacity) {
                free(states);
                return -1;
            }
            job.out = packet + pos;
            job.crc_table = g_crc;
            pos += context_model == 2 ? mdvt_ffv1::encode_slice<2>(job) : (context_model ? mdvt_ffv1::encode_slice<1>(job) : mdvt_ffv1::encode_slice<0>(job));
        }
    free(states);
    return pos;
}

// Decodes one packet (all nh x nv slices of a key frame written with this coder's parameters) into `frame` (u8x3).
// Returns 0, or the first negative slice code (see decode_slice) / -10 when the footers do not add up.
extern "C" int ffv1_host_decode_frame(const uint8_t *packet, long long packet_len, uint8_t *frame, long long row_pitch, int width,
                                      int height, int nh, int nv, int n_planes, int bgr_order, int context_model, const uint8_t *headers,
                                      const int32_t *header_len) {
    const int S = nh * nv;
    long long *starts = (long long *)malloc(sizeof(long long) * (S + 1));
    long long p = packet_len;
    starts[S] = packet_len;
    for (int s = S - 1; s >= 0; --s) {
        if (p < mdvt_ffv1::kFooterBytes) {
            free(starts);
            return -10;
        }
        const uint8_t *foot = packet + p - mdvt_ffv1::kFooterBytes;
        const long long size = (((long long)foot[0] << 16) | ((long long)foot[1] << 8) | foot[2]) + mdvt_ffv1::kFooterBytes;
        p -= size;
        if (p < 0) {
            free(starts);
            return -10;
        }
        starts[s] = p;
    }
    if (p != 0) {
        free(starts);
        return -10;
    }
    mdvt_ffv1::VlcState *states = (mdvt_ffv1::VlcState *)malloc(sizeof(mdvt_ffv1::VlcState) * 3 * mdvt_ffv1::kContexts);
    int rc = 0;
    for (int sy = 0; sy < nv && rc == 0; ++sy)
        for (int sx = 0; sx < nh && rc == 0; ++sx) {
            const int si = sy * nh + sx;
            const int x0 = (int)((long long)sx * width / nh), x1 = (int)((long long)(sx + 1) * width / nh);
            const int y0 = (int)((long long)sy * height / nv), y1 = (int)((long long)(sy + 1) * height / nv);
            mdvt_ffv1::SliceInput in;
            in.data = packet + starts[si];
            in.size = (uint32_t)(starts[si + 1] - starts[si]);
            in.header = headers + si * mdvt_ffv1::kHeaderStride;
            in.header_len = header_len[si];
            in.frame = frame + y0 * row_pitch + 3LL * x0;
            in.row_pitch = row_pitch;
            in.w = x1 - x0;
            in.h = y1 - y0;
            in.n_planes = n_planes;
            in.ib = bgr_order ? 0 : 2;
            in.ir = bgr_order ? 2 : 0;
            in.states = states;
            init_crc();
            in.crc_table = g_crc;
            rc = context_model == 2 ? mdvt_ffv1::decode_slice<2>(in) : (context_model ? mdvt_ffv1::decode_slice<1>(in) : mdvt_ffv1::decode_slice<0>(in));
        }
    free(states);
    free(starts);
    return rc;
}
