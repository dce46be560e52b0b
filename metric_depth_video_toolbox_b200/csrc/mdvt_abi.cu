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

}  // namespace mdvt

extern "C" int mdvt_abi_version(void) { return MDVT_ABI_VERSION; }

extern "C" const char *mdvt_version(void) { return "mdvt_b200 0.1 (sm_100a)"; }

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
