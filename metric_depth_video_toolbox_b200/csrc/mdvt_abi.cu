// Library-level entry points and host-side error plumbing of libmdvt_b200.so.
#include <cstdarg>
#include <cstdio>

#include "mdvt_common.cuh"

namespace mdvt {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? MDVT_ERR_NO_DEVICE : MDVT_ERR_CUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

int check_decoder(int decoder, int bit16, bool allow_f32) {
    if (decoder == MDVT_SOURCE_F32 && allow_f32) return MDVT_OK;
    if (decoder < MDVT_DECODE_D1 || decoder > MDVT_DECODE_D3) {
        set_error("unknown decoder %d", decoder);
        return MDVT_ERR_INVALID_ARGUMENT;
    }
    if (!bit16 && decoder != MDVT_DECODE_D1) {
        set_error("the 24-bit wire format exists for decoder D1 only");
        return MDVT_ERR_UNSUPPORTED;
    }
    return MDVT_OK;
}

int check_source(const mdvt_source *s) {
    MDVT_REQUIRE(s != nullptr, "mdvt_source is NULL");
    MDVT_REQUIRE(s->width > 0 && s->height > 0, "bad frame size %dx%d", s->width, s->height);
    return check_decoder(s->decoder, s->bit16, true);
}

}  // namespace mdvt

extern "C" int mdvt_abi_version(void) { return MDVT_ABI_VERSION; }

extern "C" const char *mdvt_version(void) { return "mdvt_b200 0.2 (sm_100a)"; }

extern "C" const char *mdvt_last_error(void) { return mdvt::g_error; }

extern "C" int mdvt_device_info(int *sm_count, int *l2_bytes, int *smem_optin_bytes, int *cc_major, int *cc_minor) {
    int dev = 0;
    MDVT_CUDA_TRY(cudaGetDevice(&dev));
    struct {
        int *dst;
        cudaDeviceAttr attr;
    } q[] = {{sm_count, cudaDevAttrMultiProcessorCount},
             {l2_bytes, cudaDevAttrL2CacheSize},
             {smem_optin_bytes, cudaDevAttrMaxSharedMemoryPerBlockOptin},
             {cc_major, cudaDevAttrComputeCapabilityMajor},
             {cc_minor, cudaDevAttrComputeCapabilityMinor}};
    for (auto &e : q)
        if (e.dst) MDVT_CUDA_TRY(cudaDeviceGetAttribute(e.dst, e.attr, dev));
    return MDVT_OK;
}
