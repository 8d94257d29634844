// abi.cu -- library-level entry points of the C ABI: version, thread-local error text and
// the launch counter bench.py reports as gpu_launches.
#include "common.cuh"
#include <atomic>
#include <string>

namespace evfly {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int device_sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int v = cache[dev & 63].load(std::memory_order_relaxed);
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = kNumSMs;
        cache[dev & 63].store(v, std::memory_order_relaxed);
    }
    return v;
}

}  // namespace evfly

extern "C" int evfly_abi_version(void) { return EVFLY_ABI_VERSION; }
extern "C" const char* evfly_last_error(void) { return evfly::g_err; }
extern "C" int64_t evfly_launch_count(void) { return evfly::g_launches.load(std::memory_order_relaxed); }
