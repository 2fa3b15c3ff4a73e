// cmh_core.cu — error reporting and device queries of libcmh.so.
#include "cmh_common.cuh"

#include <stdlib.h>
#include <string.h>

#include <atomic>

namespace cmh {
namespace {
thread_local char g_err[1024] = "";
std::atomic<unsigned long long> g_launches{0};
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("CMH_NO_PDL");
        return !(e && e[0] && e[0] != '0');
    }();
    return on;
}

int sm_count_cached() {
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        (void)cudaGetLastError();
        return 148;  // B200; only reached when no device is visible (CPU-side planning in tests)
    }
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            (void)cudaGetLastError();
            n = 148;
        }
        cached[dev] = n;
    }
    return cached[dev];
}
}  // namespace cmh

extern "C" {

int cmh_abi_version(void) { return CMH_ABI_VERSION; }

const char* cmh_last_error(void) { return cmh::g_err; }

unsigned long long cmh_launch_count(void) { return cmh::g_launches.load(std::memory_order_relaxed); }

int cmh_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    CMH_CUDA_TRY(cudaGetDevice(&dev));
    int sm = 0, ma = 0, mi = 0;
    CMH_CUDA_TRY(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    CMH_CUDA_TRY(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
    CMH_CUDA_TRY(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sm;
    if (cc_major) *cc_major = ma;
    if (cc_minor) *cc_minor = mi;
    return CMH_OK;
}
}
