// Microbenchmark (exploration for the accumulation kernel): throughput of
//  (a) shared-memory integer atomics on random addresses (u32, u64) at 1024 threads/CTA, 1 CTA/SM
//  (b) global RED variants on an L2-resident 8.6 MB frame (scalar s32, v4.f32) with different CTA shapes
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <typename T>
__global__ void __launch_bounds__(1024, 1) k_smem_atomics(T* out, int iters, int cells) {
    extern __shared__ unsigned char raw[];
    T* s = reinterpret_cast<T*>(raw);
    for (int i = threadIdx.x; i < cells; i += blockDim.x) s[i] = 0;
    __syncthreads();
    uint32_t h = hash32(blockIdx.x * 1024 + threadIdx.x + 1);
    for (int i = 0; i < iters; ++i) {
        h = hash32(h + i);
        atomicAdd(&s[h % cells], (T)1);
    }
    __syncthreads();
    T acc = 0;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) acc += s[i];
    if (acc == (T)0xdeadbeef) out[0] = acc;
}

__global__ void k_red_s32(const uint4* ev, long n, int* frame, int cells) {
    long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint4 r = ev[i];
        atomicAdd(frame + (r.x % cells), 1);
    }
}
__global__ void k_red_v4(const uint4* ev, long n, float4* frame, int cells) {
    long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint4 r = ev[i];
        atomicAdd(frame + (r.x % cells), make_float4(1.f, 0.f, 0.5f, 0.f));
    }
}
__global__ void k_fill(uint4* ev, long n) {
    long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) ev[i] = make_uint4(hash32((uint32_t)i), 0, 0, 0);
}

template <typename F> float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    int sms = 148;
    void* out; cudaMalloc(&out, 64);
    for (int cells_kb : {32, 128, 200}) {
        int iters = 2000;
        cudaFuncSetAttribute(k_smem_atomics<unsigned int>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_smem_atomics<unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        int c32 = cells_kb * 1024 / 4, c64 = cells_kb * 1024 / 8;
        float t32 = time_ms([&] { k_smem_atomics<unsigned int><<<sms, 1024, cells_kb * 1024>>>((unsigned int*)out, iters, c32); });
        float t64 = time_ms([&] { k_smem_atomics<unsigned long long><<<sms, 1024, cells_kb * 1024>>>((unsigned long long*)out, iters, c64); });
        double ops = (double)sms * 1024 * iters;
        printf("{\"smem_kb\": %d, \"u32_Gatomics_s\": %.1f, \"u64_Gatomics_s\": %.1f, \"u32_per_clk_per_sm\": %.2f}\n", cells_kb, ops / t32 / 1e6, ops / t64 / 1e6,
               ops / t32 / 1e6 / 148 / 1.965);
    }
    long n = 10000000;
    uint4* ev; cudaMalloc(&ev, n * 16);
    k_fill<<<1184, 256>>>(ev, n);
    int cells = 480 * 640 * 2;
    void* frame; cudaMalloc(&frame, (size_t)cells * 16);
    cudaMemset(frame, 0, (size_t)cells * 16);
    for (int threads : {128, 256, 512}) for (int per_sm : {4, 8, 16}) {
        if (threads * per_sm > 2048) continue;
        float a = time_ms([&] { k_red_s32<<<sms * per_sm, threads>>>(ev, n, (int*)frame, cells); });
        float b = time_ms([&] { k_red_v4<<<sms * per_sm, threads>>>(ev, n, (float4*)frame, cells); });
        printf("{\"threads\": %d, \"ctas_per_sm\": %d, \"red_s32_us\": %.1f, \"red_s32_G_s\": %.1f, \"red_v4_us\": %.1f, \"red_v4_G_s\": %.1f}\n", threads, per_sm, a * 1e3,
               n / a / 1e6, b * 1e3, n / b / 1e6);
    }
    return 0;
}
