// tc_stem.cu -- the UNet stem (learner_models.py OrigUNet unet_e11: Conv2d(Cin, 32, 3) + ReLU, Cin = 1|2) on tcgen05.
//
// On CUDA cores this layer is FMA-bound (18 x 32 FMAs per pixel: 0.58 ms of issue time for 400 frames) although it
// only has to write 64 bytes per pixel. As a GEMM it is [pixels x 32] = im2col[pixels x 18->32] . W^T[32 x 32]: each
// thread builds the im2col row of its pixel (bf16, K zero-padded to 32) directly in shared memory in the 64-byte
// swizzled K-major layout the tensor core reads (16-byte chunk c of row r lives at chunk c ^ ((r >> 1) & 3)), one
// thread issues two tcgen05.mma (M128 N32 K16) into a 32-column TMEM accumulator, and every thread reads its own
// pixel's 32 outputs back (tcgen05.ld 32x32b), adds the fp32 bias, applies ReLU and stores 64 bytes of bf16 NHWC.
// No TMA: the A operand never exists in global memory. Many small CTAs per SM hide the per-tile latency chain.
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"

namespace evfly {

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

template <int CIN>
__global__ void __launch_bounds__(128)
k_stem_tc(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, uint4* __restrict__ out,
          int H, int W, long long total_px, int n_tiles) {
    __shared__ __align__(1024) uint8_t s_a[128 * 64];     // im2col tile, [128 pixels][32 k] bf16, 64B swizzle
    __shared__ __align__(1024) uint8_t s_b[32 * 64];      // weights [32 cout][32 k] bf16, 64B swizzle
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[32];
    constexpr int K_REAL = CIN * 9;
    const int t = threadIdx.x, warp = t >> 5;
    if (warp == 0) tmem_alloc(&s_tmem, 32);
    if (t == 0) {
        mbar_init(&s_bar, 1);
        fence_barrier_init();
    }
    if (t < 32) s_bias[t] = bias[t];
    {   // weights: PyTorch [32][CIN][3][3] is already k = ci*9 + tap contiguous per output channel
        const int row = t >> 2, c = t & 3;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = c * 8 + e;
            f[e] = k < K_REAL ? w[row * K_REAL + k] : 0.f;
        }
        uint4 pk;
        pk.x = pack_bf16x2(f[0], f[1]); pk.y = pack_bf16x2(f[2], f[3]); pk.z = pack_bf16x2(f[4], f[5]); pk.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(s_b + row * 64 + ((c ^ ((row >> 1) & 3)) << 4)) = pk;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const uint32_t a_base = smem_u32(s_a), b_base = smem_u32(s_b);
    constexpr uint32_t idesc = make_idesc_bf16(128, 32);
    const long long plane = (long long)H * W;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long i = (long long)tile * 128 + t;
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;
        if (i < total_px) {
            // 32-bit division when the pixel count allows (always, in practice): the 64-bit one costs ~100 instructions per pixel
            const long long n = total_px < (1ll << 31) ? (long long)((unsigned)i / (unsigned)plane) : i / plane;
            const int rem = (int)(i - n * plane);
            const int oh = rem / W, ow = rem - oh * W;
            if (oh < H - 2 && ow < W - 2) {
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) {
                    const float* xp = x + (n * CIN + ci) * plane + rem;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) v[ci * 9 + tap] = __ldg(xp + (tap / 3) * W + (tap % 3));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4 pk;
            pk.x = pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]); pk.y = pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]);
            pk.z = pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]); pk.w = pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]);
            *reinterpret_cast<uint4*>(s_a + t * 64 + ((c ^ ((t >> 1) & 3)) << 4)) = pk;
        }
        fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 2; ++k)
                umma_bf16(tmem_base, make_smem_desc(a_base + k * 32, 512, kLayoutSw64), make_smem_desc(b_base + k * 32, 512, kLayoutSw64), idesc, k > 0);
            umma_commit(&s_bar);
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), r);
        tmem_ld_wait();
        // bias + ReLU -> bf16, staged through this thread's (now free) row of the A tile so that the global stores
        // are contiguous: store instruction q of a warp writes pixels 8q..8q+7 = 512 consecutive bytes
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(r[q * 8 + e]) + s_bias[q * 8 + e], 0.f);
            uint4 pk;
            pk.x = pack_bf16x2(f[0], f[1]); pk.y = pack_bf16x2(f[2], f[3]); pk.z = pack_bf16x2(f[4], f[5]); pk.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(s_a + t * 64 + ((q ^ ((t >> 1) & 3)) << 4)) = pk;
        }
        __syncwarp();
        {
            const int lane = t & 31;
            const long long px0 = (long long)tile * 128 + warp * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int row = warp * 32 + q * 8 + (lane >> 2), c = lane & 3;
                const uint4 pk = *reinterpret_cast<const uint4*>(s_a + row * 64 + ((c ^ ((row >> 1) & 3)) << 4));
                const long long px = px0 + q * 8 + (lane >> 2);
                if (px < total_px) out[px * 4 + c] = pk;
            }
        }
        tc_fence_before();
        __syncthreads();               // the accumulator and the A tile are free again
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 32);
    }
}

}  // namespace evfly

using namespace evfly;

extern "C" int evfly_stem_conv3x3_bf16(const float* d_x, const float* d_w, const float* d_bias, void* d_out, int N, int Cin,
                                       int H, int W, void* stream) {
    EVFLY_REQUIRE(d_x && d_w && d_bias && d_out && N >= 0 && (Cin == 1 || Cin == 2) && H >= 3 && W >= 3, "stem_conv3x3_bf16: bad argument (Cin must be 1 or 2)");
    if (N == 0) return EVFLY_OK;
    const long long total = (long long)N * H * W;
    const long long tiles = (total + 127) / 128;
    EVFLY_REQUIRE(tiles < (1ll << 31), "stem_conv3x3_bf16: too many pixels");
    // persistent tile loop: one wave of small CTAs (56 registers x 128 threads -> 9 resident per SM; the sweep in
    // scripts/bench_stem.py is flat from 9 up)
    static const char* env = getenv("EVFLY_STEM_CTAS_PER_SM");      // tuning knob
    const long long wave = (long long)kNumSMs * (env ? atoi(env) : 16);
    const int grid = (int)(tiles < wave ? tiles : wave);
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 1)
        k_stem_tc<1><<<grid, 128, 0, st>>>(d_x, d_w, d_bias, reinterpret_cast<uint4*>(d_out), H, W, total, (int)tiles);
    else
        k_stem_tc<2><<<grid, 128, 0, st>>>(d_x, d_w, d_bias, reinterpret_cast<uint4*>(d_out), H, W, total, (int)tiles);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}
