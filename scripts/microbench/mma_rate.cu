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

// variant 3: the per-tile pattern of the halo convs -- `per_tile` MMAs into one accumulator (the first overwrites), then
// `commits` tcgen05.commit to barriers nobody waits on, accumulators rotating over `nacc`. Question: does a commit (or the
// accumulator switch) cost tensor-pipe time beyond the MMAs themselves?
__global__ void __launch_bounds__(128) k_tile_pattern(int N, int per_tile, int commits, int first_overwrites, int nacc, int tiles, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t sink[2];
    __shared__ uint32_t s_tmem;
    if (threadIdx.x < 32) tmem_alloc(&s_tmem, 512);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&sink[0], 1 << 20); mbar_init(&sink[1], 1 << 20); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = s_tmem;
    if (threadIdx.x < 32) {
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t da = make_smem_desc(a, 10 * 64, kLayoutSw64), db = make_smem_desc(b, 8 * 64, kLayoutSw64);
        fence_proxy_async();
        if (elect_one()) {
            for (int i = 0; i < 8; ++i) umma_bf16(tm, da, db, idesc, i > 0);
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t0 = clock64();
        int acc = 0;
        for (int t = 0; t < tiles; ++t) {
            if (elect_one()) {
                for (int i = 0; i < per_tile; ++i) umma_bf16(tm + (uint32_t)(acc * N), da + (uint64_t)((i & 7) * 4), db + (uint64_t)((i & 7) * 4), idesc, (i > 0) || !first_overwrites);
                if (commits >= 1) umma_commit(&sink[0]);
                if (commits >= 2) umma_commit(&sink[1]);
            }
            __syncwarp();
            if (++acc == nacc) acc = 0;
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

// variant 4: what is the fixed ~300 clk per tile of variant 3? MODE bits: 1 = MMAs unrolled (18 per tile, constant descriptor offsets),
// 2 = no __syncwarp between tiles, 4 = predicated issue from convergent code (no divergent region), 8 = two issuing warps (tiles
// split between them, separate accumulators), 16 = per tile also wait on an already completed mbarrier and fence (the real kernel's
// full/tempty waits), 32 = one commit per tile
template <int MODE>
__global__ void __launch_bounds__(128) k_tile_pattern2(int tiles, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ __align__(8) uint64_t sink[2];
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t s_tmem;
    constexpr int N = 32;
    if (threadIdx.x < 32) tmem_alloc(&s_tmem, 512);
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&sink[0], 1 << 20); mbar_init(&sink[1], 1 << 20); mbar_init(&done_bar, 1); fence_barrier_init(); mbar_arrive(&done_bar); /* phase 0 complete: waits on parity 0 pass */ }
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = s_tmem;
    const int warp = threadIdx.x >> 5;
    constexpr int kWarps = (MODE & 8) ? 2 : 1;
    if (warp < kWarps) {
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
        constexpr uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t da = make_smem_desc(a, 10 * 64, kLayoutSw64), db = make_smem_desc(b, 8 * 64, kLayoutSw64);
        fence_proxy_async();
        const uint32_t elected = elect_one() ? 1u : 0u;
        const long long t0 = clock64();
        int acc = 0;
        for (int t = warp; t < tiles; t += kWarps) {
            if (MODE & 16) {
                mbar_wait(&done_bar, 0);           // already complete
                tc_fence_after();
            }
            const uint32_t tmem_d = tm + (uint32_t)((warp * 4 + acc) * N);
            if (MODE & 4) {
#pragma unroll
                for (int i = 0; i < 18; ++i) umma_bf16_pred(tmem_d, da + (uint64_t)(((i / 2) * 11 * 64 + (i & 1) * 32) >> 4), db + (uint64_t)((i * 1024) >> 4), idesc, i > 0, elected);
                if (MODE & 32) umma_commit_pred(&sink[0], elected);
            } else {
                if (elect_one()) {
                    if (MODE & 1) {
#pragma unroll
                        for (int i = 0; i < 18; ++i) umma_bf16(tmem_d, da + (uint64_t)(((i / 2) * 11 * 64 + (i & 1) * 32) >> 4), db + (uint64_t)((i * 1024) >> 4), idesc, i > 0);
                    } else {
#pragma unroll 1
                        for (int i = 0; i < 18; ++i) umma_bf16(tmem_d, da + (uint64_t)((i & 7) * 4), db + (uint64_t)((i & 7) * 4), idesc, i > 0);
                    }
                    if (MODE & 32) umma_commit(&sink[0]);
                }
                if (!(MODE & 2)) __syncwarp();
            }
            if (++acc == 4) acc = 0;
        }
        if (elect_one()) umma_commit(&bar[warp]);
        __syncwarp();
        mbar_wait(&bar[warp], 0);
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int MODE>
static void run_pattern2(long long* d_out) {
    cudaFuncSetAttribute(k_tile_pattern2<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    const int tiles = 1024;
    long long c[2] = {0, 0};
    cudaMemset(d_out, 0, 16);
    k_tile_pattern2<MODE><<<1, 128, 68 * 1024>>>(tiles, d_out);
    cudaError_t e = cudaMemcpy(c, d_out, 16, cudaMemcpyDeviceToHost);
    const long long mx = c[0] > c[1] ? c[0] : c[1];
    printf("{\"pattern\": \"tile2\", \"unrolled\": %d, \"no_syncwarp\": %d, \"predicated_issue\": %d, \"issuing_warps\": %d, \"wait_and_fence\": %d, \"commit\": %d, \"cycles_per_tile\": %.1f, \"err\": \"%s\"}\n",
           MODE & 1, (MODE >> 1) & 1, (MODE >> 2) & 1, (MODE & 8) ? 2 : 1, (MODE >> 4) & 1, (MODE >> 5) & 1, (double)mx / tiles, cudaGetErrorString(e));
    fflush(stdout);
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 16);
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
    cudaFuncSetAttribute(k_tile_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    for (int N : {32, 64})
        for (int per_tile : {18, 36})
            for (int commits : {0, 1, 2})
                for (int fo : {0, 1})
                    for (int nacc : {1, 4}) {
                        long long c = 0;
                        const int tiles = 512;
                        k_tile_pattern<<<1, 128, 68 * 1024>>>(N, per_tile, commits, fo, nacc, tiles, d_out);
                        cudaError_t e = cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
                        printf("{\"pattern\": \"tile\", \"N\": %d, \"mmas_per_tile\": %d, \"commits_per_tile\": %d, \"first_overwrites\": %d, \"accumulators\": %d, \"cycles_per_tile\": %.1f, \"cycles_per_mma\": %.1f, \"err\": \"%s\"}\n",
                               N, per_tile, commits, fo, nacc, (double)c / tiles, (double)c / tiles / per_tile, cudaGetErrorString(e));
                    }
    run_pattern2<0>(d_out); run_pattern2<1>(d_out); run_pattern2<3>(d_out); run_pattern2<4>(d_out); run_pattern2<8>(d_out); run_pattern2<9>(d_out);
    run_pattern2<12>(d_out); run_pattern2<16>(d_out); run_pattern2<17>(d_out); run_pattern2<20>(d_out); run_pattern2<33>(d_out); run_pattern2<36>(d_out);
    run_pattern2<49>(d_out); run_pattern2<52>(d_out); run_pattern2<60>(d_out); run_pattern2<57>(d_out);
    return 0;
}