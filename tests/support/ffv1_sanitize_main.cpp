// Test support only: the shared slice coder text (csrc/mdvt_ffv1_slice.h, host build) under AddressSanitizer / UBSan:
// encode -> decode round trips over tiny and ragged slices, both context models, with and without alpha, and decoding of
// bit-flipped packets (must stay inside the packet and the frame).  argv[1] = path of libmdvt_b200.so (stream setup).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <dlfcn.h>
#include <stdint.h>
extern "C" long long ffv1_host_encode_frame(const uint8_t *, long long, int, int, int, int, int, int, int, const uint8_t *, const int32_t *, uint8_t *, long long);
extern "C" int ffv1_host_decode_frame(const uint8_t *, long long, uint8_t *, long long, int, int, int, int, int, int, int, const uint8_t *, const int32_t *);
extern "C" long long ffv1_host_slice_capacity(int, int, int);
typedef int (*setup_t)(int, int, int, int, int, int, uint8_t *, int, int *, uint8_t *, int32_t *);
int main(int argc, char **argv) {
    if (argc < 2) return 2;
    void *h = dlopen(argv[1], RTLD_NOW);
    if (!h) { printf("dlopen: %s\n", dlerror()); return 2; }
    setup_t setup = (setup_t)dlsym(h, "mdvt_ffv1_stream_setup");
    struct { int w, h, nh, nv; } cases[] = {{1, 1, 1, 1}, {2, 1, 2, 1}, {1, 5, 1, 5}, {7, 3, 7, 3}, {7, 3, 1, 1}, {33, 17, 4, 3}, {64, 48, 5, 7}, {130, 70, 16, 9}, {97, 61, 3, 2}};
    unsigned seed = 12345;
    int bad = 0;
    for (auto &c : cases)
        for (int alpha = 0; alpha < 2; ++alpha)
            for (int model = 0; model < 2; ++model)
                for (int kind = 0; kind < 4; ++kind) {
                    const int S = c.nh * c.nv;
                    std::vector<uint8_t> frame((size_t)c.w * c.h * 3), out((size_t)c.w * c.h * 3, 0xA5), headers((size_t)S * 16);
                    std::vector<int32_t> lens(S);
                    for (auto &b : frame) {
                        seed = seed * 1664525u + 1013904223u;
                        b = kind == 0 ? (uint8_t)(seed >> 24) : kind == 1 ? 77 : kind == 2 ? (uint8_t)(((seed >> 24) & 1) * 255) : (uint8_t)(128 + ((seed >> 28) & 3));
                    }
                    uint8_t cfg[64];
                    int cfg_len = 0;
                    if (setup(c.w, c.h, c.nh, c.nv, alpha, model, cfg, 64, &cfg_len, headers.data(), lens.data())) { printf("setup failed\n"); return 3; }
                    long long cap = 0;
                    for (int s = 0; s < S; ++s) cap += ffv1_host_slice_capacity((c.w + c.nh - 1) / c.nh + 1, (c.h + c.nv - 1) / c.nv + 1, 3 + alpha);
                    std::vector<uint8_t> packet((size_t)cap);   // exact worst-case allocation: ASan guards the end
                    long long n = ffv1_host_encode_frame(frame.data(), (long long)c.w * 3, c.w, c.h, c.nh, c.nv, 3 + alpha, kind & 1, model, headers.data(), lens.data(), packet.data(), cap);
                    if (n <= 0) { printf("encode failed %dx%d\n", c.w, c.h); ++bad; continue; }
                    std::vector<uint8_t> exact(packet.begin(), packet.begin() + n);   // exact-size copy: reads past the packet trip ASan
                    int rc = ffv1_host_decode_frame(exact.data(), n, out.data(), (long long)c.w * 3, c.w, c.h, c.nh, c.nv, 3 + alpha, kind & 1, model, headers.data(), lens.data());
                    if (rc != 0 || memcmp(out.data(), frame.data(), frame.size()) != 0) { printf("MISMATCH %dx%d %dx%d alpha %d model %d kind %d rc %d\n", c.w, c.h, c.nh, c.nv, alpha, model, kind, rc); ++bad; }
                    // corrupted packets must not read or write out of bounds
                    for (int t = 0; t < 6; ++t) {
                        std::vector<uint8_t> dmg(exact);
                        seed = seed * 1664525u + 1013904223u;
                        dmg[seed % dmg.size()] ^= (uint8_t)(1u << ((seed >> 8) & 7));
                        ffv1_host_decode_frame(dmg.data(), n, out.data(), (long long)c.w * 3, c.w, c.h, c.nh, c.nv, 3 + alpha, kind & 1, model, headers.data(), lens.data());
                    }
                }
    printf("done, %d bad\n", bad);
    return bad != 0;
}
