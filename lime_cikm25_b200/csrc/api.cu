// Library-level plumbing of liblime_b200.so: error string, launch counter, device queries.
#include <stdarg.h>

#include <atomic>
#include <string.h>

#include "common.cuh"

namespace lime {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};   // process-wide: autograd runs the backward kernels on its own thread

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;   // B200
    }
    return cached;
}

}  // namespace lime

extern "C" int lime_abi_version(void) { return LIME_B200_ABI_VERSION; }

extern "C" const char *lime_last_error(void) { return lime::g_error; }

extern "C" int lime_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int64_t lime_launch_count(void) { return lime::g_launches.load(std::memory_order_relaxed); }

extern "C" void lime_launch_count_reset(void) { lime::g_launches.store(0, std::memory_order_relaxed); }

extern "C" int64_t lime_sizeof_news_cache(void) { return (int64_t)sizeof(LimeNewsCache); }

extern "C" int64_t lime_sizeof_impressions(void) { return (int64_t)sizeof(LimeImpressions); }
