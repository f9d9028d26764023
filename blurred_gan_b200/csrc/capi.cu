// ABI bookkeeping shared by every translation unit of libbgan_b200.so.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace bgan {

namespace {
thread_local char g_error[512] = "";
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int device_sm_count() {
    thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || cached <= 0) cached = 148;
        cached_dev = dev;
    }
    return cached;
}

int device_max_smem_optin() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 227 * 1024;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0) v = 227 * 1024;
    return v;
}

}  // namespace bgan

extern "C" int bgan_version(void) { return BGAN_ABI_VERSION; }

extern "C" const char* bgan_last_error(void) { return bgan::g_error; }
