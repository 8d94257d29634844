// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, bf16) as a function of N and of the
// shared-memory layout of the operands (128/64/32-byte swizzle). Question: how much of a small-N MMA is the A-operand
// read, and does a narrower swizzle row make it cheaper? Data is garbage; only the issue-to-completion time is read.
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include "../../evfly_b200/csrc/tc_common.cuh"
using namespace evfly;

// variant 1: the WHOLE warp 0 runs the issue loop convergently, one elected lane issues (no divergent region)
__global__ void __launch_bounds__(128) k_rate_elect(int N, uint32_t layout, uint32_t row_bytes, int iters, long long* out, int shift_rows = 0, int sbo_rows = 8) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    if (threadIdx.x < 32) tmem_alloc(&s_tmem, 512);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = s_tmem;
    if (threadIdx.x < 32) {
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t da = make_smem_desc(a + shift_rows * row_bytes, sbo_rows * row_bytes, layout), db = make_smem_desc(b, 8 * row_bytes, layout);
        fence_proxy_async();
        if (elect_one()) {
            for (int i = 0; i < 8; ++i) umma_bf16(tm, da, db, idesc, i > 0);
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < iters; ++i) {
            // descriptors vary per MMA like in a real k-loop (start address + 32 B per k-step, wrapping)
            const uint64_t dai = da + (uint64_t)((i & 3) * 2), dbi = db + (uint64_t)((i & 3) * 2);
            if (elect_one()) umma_bf16(tm + (uint32_t)((i & 1) * 256), dai, dbi, idesc, 1);
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 1);
        const long long t1 = clock64();
        if (threadIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

__global__ void __launch_bounds__(128) k_rate(int N, uint32_t layout, uint32_t row_bytes, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    if (threadIdx.x < 32) tmem_alloc(&s_tmem, 512);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = s_tmem;
    if (threadIdx.x == 0) {
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t da = make_smem_desc(a, 8 * row_bytes, layout), db = make_smem_desc(b, 8 * row_bytes, layout);
        fence_proxy_async();
        // warm
        for (int i = 0; i < 8; ++i) umma_bf16(tm, da, db, idesc, i > 0);
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) umma_bf16(tm + (uint32_t)((i & 1) * 256), da, db, idesc, 1);
        umma_commit(&bar);
        mbar_wait(&bar, 1);
        const long long t1 = clock64();
        out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaFuncSetAttribute(k_rate_elect, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    const int iters = 4096;
    struct { const char* name; uint32_t layout, row; } L[] = {{"sw128", 2, 128}, {"sw64", 4, 64}};
    for (auto& l : L)
        for (int N : {8, 16, 32, 64, 128, 256}) {
            k_rate<<<1, 128, 68 * 1024>>>(N, l.layout, l.row, iters, d_out);
            long long c = 0, c2 = 0;
            cudaError_t e = cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
            k_rate_elect<<<1, 128, 68 * 1024>>>(N, l.layout, l.row, iters, d_out);
            cudaError_t e2 = cudaMemcpy(&c2, d_out, 8, cudaMemcpyDeviceToHost);
            printf("{\"layout\": \"%s\", \"N\": %d, \"cycles_per_mma_one_thread\": %.1f, \"cycles_per_mma_elect_warp\": %.1f, \"err\": \"%s/%s\"}\n", l.name, N,
                   (double)c / iters, (double)c2 / iters, cudaGetErrorString(e), cudaGetErrorString(e2));
        }
    // halo-style A operand: descriptor starts `shift` rows into the tile, 8-row groups 10 rows apart (tc_conv_halo.cu)
    for (auto& l : L)
        for (int shift : {0, 1, 2, 3, 10, 11, 21, 22}) {
            long long c2 = 0;
            k_rate_elect<<<1, 128, 68 * 1024>>>(32, l.layout, l.row, iters, d_out, shift, 10);
            cudaError_t e2 = cudaMemcpy(&c2, d_out, 8, cudaMemcpyDeviceToHost);
            printf("{\"layout\": \"%s\", \"N\": 32, \"halo_shift_rows\": %d, \"sbo_rows\": 10, \"cycles_per_mma_elect_warp\": %.1f, \"err\": \"%s\"}\n", l.name, shift, (double)c2 / iters, cudaGetErrorString(e2));
        }
    return 0;
}
