// common.cuh -- shared host/device helpers of libevfly_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/evfly_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libevfly_b200 is written for sm_100a (B200) only"
#endif

namespace evfly {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- error reporting / launch accounting (defined in abi.cu) ---------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define EVFLY_REQUIRE(cond, ...)                                  \
    do {                                                          \
        if (!(cond)) {                                            \
            evfly::set_error(__VA_ARGS__);                        \
            return EVFLY_ERR_INVALID_ARG;                         \
        }                                                         \
    } while (0)

#define EVFLY_CUDA(call)                                                              \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            evfly::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                     \
            return EVFLY_ERR_CUDA;                                                    \
        }                                                                             \
    } while (0)

// call after every kernel launch: catches configuration errors without synchronising
#define EVFLY_LAUNCHED()                      \
    do {                                      \
        evfly::count_launch();                \
        EVFLY_CUDA(cudaPeekAtLastError());    \
    } while (0)

// ---- device helpers -------------------------------------------------------------------
#ifdef __CUDACC__
// streaming 16-byte load that does not pollute L1 (events are read exactly once)
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
#endif

// One-time setup that belongs to a DEVICE's context (cudaFuncSetAttribute): done once per device and call site, so a
// process that drives several GPUs configures every one of them (a function-local `static bool` configured
// only the first device and the kernels of the second failed to launch with > 48 KB of shared memory).
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask{0ull};
    unsigned long long bit() const {
        int dev = 0;
        cudaGetDevice(&dev);
        return 1ull << (dev & 63);
    }
};
#define EVFLY_SMEM_ATTR(bytes, ...)                                                                                   \
    do {                                                                                                               \
        static evfly::PerDeviceOnce once__;                                                                            \
        const unsigned long long bit__ = once__.bit();                                                                 \
        if (!(once__.mask.load(std::memory_order_acquire) & bit__)) {                                                  \
            EVFLY_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));  \
            once__.mask.fetch_or(bit__, std::memory_order_release);                                                    \
        }                                                                                                              \
    } while (0)

// number of SMs of the CURRENT device (cached per device); persistent kernels with a grid barrier size their grid from it
int device_sm_count();

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// persistent-style grid for streaming kernels: `ctas_per_sm` CTAs on each of the 148 SMs,
// capped so that tiny inputs do not launch idle CTAs
inline int stream_grid(int64_t work_items, int items_per_cta, int ctas_per_sm) {
    int64_t want = ceil_div(work_items, items_per_cta);
    int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace evfly
