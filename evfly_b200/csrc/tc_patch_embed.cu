// tc_patch_embed.cu -- OverlapPatchMerging (ViTsubmodules.py:15-34: Conv2d(k, stride, pad) + LayerNorm) for the two
// token stages of the ViT on tcgen05, for batches. As a GEMM it is [tokens x Cout] = im2col[tokens x k*k*Cin] . W^T;
// the im2col row of a token never exists in global memory: each thread gathers the patch of ITS token (stage 1: 49
// fp32 pixels of the depth image; stage 2: 9 taps x 32 bf16 channels, 64 contiguous bytes per tap) straight into
// 128-byte-swizzled K-major shared memory (K zero-padded to a multiple of 64), one thread issues the MMAs
// (M128, N = Cout, K16) into TMEM, and every thread then owns the full Cout-wide row of its token: bias, LayerNorm
// (mean / variance in registers, no shuffles) and the bf16 store happen in the epilogue.
#include "common.cuh"
#include "tc_common.cuh"

namespace evfly {

__device__ __forceinline__ uint32_t pe_pack(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

// NBLK = number of 64-wide K blocks (k*k*Cin <= 64*NBLK)
template <int NBLK, int COUT, bool F32_IN>
__global__ void __launch_bounds__(128)
k_patch_embed_tc(const void* __restrict__ xin, const float* __restrict__ w_kc, const float* __restrict__ bias, const float* __restrict__ gamma,
                 const float* __restrict__ beta, __nv_bfloat16* __restrict__ out, long long tokens, int H, int W, int Cin, int k, int s, int p,
                 int OH, int OW, float eps, int n_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_a = smem;                                   // [NBLK][128 rows][128 B]
    uint8_t* s_b = smem + NBLK * 16384;                    // [NBLK][COUT rows][128 B]
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[COUT], s_gamma[COUT], s_beta[COUT];
    const int t = threadIdx.x, warp = t >> 5;
    const int K = k * k * Cin;
    if (warp == 0) tmem_alloc(&s_tmem, COUT);
    if (t == 0) {
        mbar_init(&s_bar, 1);
        fence_barrier_init();
    }
    if (t < COUT) {
        s_bias[t] = bias[t];
        s_gamma[t] = gamma[t];
        s_beta[t] = beta[t];
    }
    // weights: w_kc is fp32 [K][COUT]; B operand = [COUT rows][K] bf16, K-major, 128B swizzle, zero beyond K
    for (int i = t; i < NBLK * COUT * 8; i += 128) {
        const int c = i & 7, co = (i >> 3) % COUT, blk = i / (8 * COUT);
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int kk = blk * 64 + c * 8 + e;
            f[e] = kk < K ? w_kc[(long long)kk * COUT + co] : 0.f;
        }
        uint4 pk;
        pk.x = pe_pack(f[0], f[1]); pk.y = pe_pack(f[2], f[3]); pk.z = pe_pack(f[4], f[5]); pk.w = pe_pack(f[6], f[7]);
        *reinterpret_cast<uint4*>(s_b + blk * (COUT * 128) + co * 128 + ((c ^ (co & 7)) << 4)) = pk;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const uint32_t a_base = smem_u32(s_a), b_base = smem_u32(s_b);
    constexpr uint32_t idesc = make_idesc_bf16(128, COUT);
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long tok = (long long)tile * 128 + t;
        const bool live = tok < tokens;
        const int ow = live ? (int)(tok % OW) : 0, oh = live ? (int)((tok / OW) % OH) : 0;
        const long long b = live ? tok / ((long long)OW * OH) : 0;
        uint8_t* arow = s_a + t * 128;
        const int sw = t & 7;
        if (F32_IN) {
            // Cin == 1: K index = kh*k + kw; 64 values -> 8 chunks
            const float* img = reinterpret_cast<const float*>(xin) + b * (long long)H * W;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int kk = c * 8 + e;
                    const int kh = kk / k, kw = kk - kh * k;
                    const int ih = oh * s - p + kh, iw = ow * s - p + kw;
                    f[e] = (live && kk < K && (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) ? __ldg(img + (long long)ih * W + iw) : 0.f;
                }
                uint4 pk;
                pk.x = pe_pack(f[0], f[1]); pk.y = pe_pack(f[2], f[3]); pk.z = pe_pack(f[4], f[5]); pk.w = pe_pack(f[6], f[7]);
                *reinterpret_cast<uint4*>(arow + ((c ^ sw) << 4)) = pk;
            }
        } else {
            // bf16 NHWC input, Cin % 8 == 0: K index = tap*Cin + ci; chunk j of the row = 8 consecutive ci of one tap
            const __nv_bfloat16* img = reinterpret_cast<const __nv_bfloat16*>(xin) + b * (long long)H * W * Cin;
            const int cpt = Cin >> 3;                          // chunks per tap
            for (int j = 0; j < NBLK * 8; ++j) {
                const int tap = j / cpt, cc = j - tap * cpt;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (live && tap < k * k) {
                    const int kh = tap / k, kw = tap - kh * k;
                    const int ih = oh * s - p + kh, iw = ow * s - p + kw;
                    if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W)
                        v = *reinterpret_cast<const uint4*>(img + ((long long)ih * W + iw) * Cin + cc * 8);
                }
                *reinterpret_cast<uint4*>(arow + (j >> 3) * 16384 + (((j & 7) ^ sw) << 4)) = v;
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk)
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                    umma_bf16(tmem_base, make_smem_desc(a_base + blk * 16384 + k4 * 32, 1024, kLayoutSw128),
                              make_smem_desc(b_base + blk * (COUT * 128) + k4 * 32, 1024, kLayoutSw128), idesc, (blk | k4) != 0);
            umma_commit(&s_bar);
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        tc_fence_after();
        float v[COUT];
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[c0 + e] = __uint_as_float(r[e]) + s_bias[c0 + e];
        }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < COUT; ++c) sum += v[c];
        const float mean = sum / (float)COUT;
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
            v[c] -= mean;
            var = fmaf(v[c], v[c], var);
        }
        const float rstd = rsqrtf(var / (float)COUT + eps);
        if (live) {
            uint4* o = reinterpret_cast<uint4*>(out + tok * COUT);
#pragma unroll
            for (int q = 0; q < COUT / 8; ++q) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = v[q * 8 + e] * rstd * s_gamma[q * 8 + e] + s_beta[q * 8 + e];
                uint4 pk;
                pk.x = pe_pack(f[0], f[1]); pk.y = pe_pack(f[2], f[3]); pk.z = pe_pack(f[4], f[5]); pk.w = pe_pack(f[6], f[7]);
                o[q] = pk;
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, COUT);
    }
}

template <int NBLK, int COUT, bool F32_IN>
static int launch_pe(const void* x, const float* w, const float* bias, const float* gamma, const float* beta, void* out, long long tokens, int H,
                     int W, int Cin, int k, int s, int p, int OH, int OW, float eps, cudaStream_t st) {
    constexpr int smem = NBLK * 16384 + NBLK * COUT * 128 + 1024;
    EVFLY_SMEM_ATTR(smem, k_patch_embed_tc<NBLK, COUT, F32_IN>);
    const long long tiles = (tokens + 127) / 128;
    const int per_sm = smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 8);
    const long long wave = (long long)kNumSMs * per_sm;
    const int grid = (int)(tiles < wave ? tiles : wave);
    k_patch_embed_tc<NBLK, COUT, F32_IN><<<grid, 128, smem, st>>>(x, w, bias, gamma, beta, reinterpret_cast<__nv_bfloat16*>(out), tokens, H, W, Cin, k,
                                                                  s, p, OH, OW, eps, (int)tiles);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

// returns -1 when the shape is not one the tensor-core kernel is instantiated for
int patch_embed_tc_dispatch(const void* x, int x_is_f32, const float* w, const float* bias, const float* gamma, const float* beta, void* out,
                            long long tokens, int H, int W, int Cin, int Cout, int k, int s, int p, int OH, int OW, float eps, cudaStream_t st) {
    const int K = k * k * Cin;
    if (x_is_f32 && Cin == 1 && K <= 64 && Cout == 32)
        return launch_pe<1, 32, true>(x, w, bias, gamma, beta, out, tokens, H, W, Cin, k, s, p, OH, OW, eps, st);
    if (!x_is_f32 && Cin % 8 == 0 && K <= 320 && K > 256 && Cout == 64)
        return launch_pe<5, 64, false>(x, w, bias, gamma, beta, out, tokens, H, W, Cin, k, s, p, OH, OW, eps, st);
    return -1;
}

}  // namespace evfly
