// vit_fused.cu -- one CTA per sample: a whole MixFFN block of a ViT stage on the tensor cores with the expanded
// activation resident on chip (learner/ViTsubmodules.py:85-120 MixFFN + the residual add and LayerNorm of
// MixTransformerEncoderLayer.forward :143-146):
//
//     x -> mlp1 (C -> 8C) -> grouped 3x3 conv (groups = C, 8 -> 8 per group, padding 'same') -> exact GELU
//       -> mlp2 (8C -> C) -> + x -> LayerNorm(C)
//
// Round 1 ran this as four launches (two tcgen05 GEMMs, a CUDA-core grouped conv, a LayerNorm) with the 8C-wide
// activation going through HBM/L2 twice; here nothing but the tokens (C = 32 / 64 bf16 per token) leaves the SM.
//
// Layout. The sample's tokens sit in shared memory on a PADDED grid: row g = (y+1)*WP + (x+1), WP = W + 1 (one
// zero column doubles as the right pad of a row and the left pad of the next; a zero row above and below), as a
// K-major UMMA operand (64- or 128-byte swizzle applied by the writing threads on absolute address bits, like TMA
// would). On that grid the 3x3 'same' conv of a 32-channel slice of the expanded activation is nine SHIFTED GEMMs:
// A = the slice read through a descriptor that starts (dy*WP + dx) rows away, B = a 32x32 block-diagonal matrix of
// the four 8x8 groups of the slice (the tensor core multiplies the zeros too: 4x the algorithmic MACs, but a 32-wide
// MMA costs the same 40 clocks as anything narrower, so a denser packing would not run faster).
// Per 32-channel slice (8 slices in stage 1, 16 in stage 2), M tiles of 128 grid rows:
//     mlp1 slice  [M x C]*[C x 32]   -> TMEM P -> +bias, zero at the pad positions, bf16 -> shared Y   (A of the conv)
//     conv        9 x [M x 32]*[32 x 32] shifted -> TMEM P -> +bias, GELU, bf16 -> shared Y (in place: A of mlp2)
//     mlp2 slice  [M x 32]*[32 x C]  -> TMEM Q, accumulated over the slices
// then Q + bias + residual -> LayerNorm -> bf16 tokens. Weights of a slice arrive as one pre-swizzled image by a 1-D
// bulk copy (cp.async.bulk, mbarrier), double buffered so that slice j+1 loads while slice j computes.
#include "tc_common.cuh"

namespace evfly {

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

template <int C, int TH, int TW>
struct FfnCfg {
    static constexpr int CE = 8 * C;
    static constexpr int SLICE = 32;                       // expanded channels per slice = 4 groups
    static constexpr int NSLICE = CE / SLICE;
    static constexpr int WP = TW + 1;
    static constexpr int G_FIRST = WP + 1, G_LAST = TH * WP + TW;
    static constexpr int G0 = (G_FIRST / 8) * 8;
    static constexpr int MT = (G_LAST - G0 + 1 + 127) / 128;
    static constexpr int LEAD = 8 + ((WP + 1 > G0) ? ((WP + 1 - G0 + 7) / 8) * 8 : 0);   // rows in front of grid row 0
    static constexpr int ROWS = ((LEAD + G0 + MT * 128 + WP + 1 + 15) / 16) * 16;        // x 64 bytes = a multiple of 1024
    static constexpr int XROWB = C * 2;                    // bytes per token row of the X operand (64: SW64, 128: SW128)
    static constexpr int YROWB = SLICE * 2;                // 64: SW64
    static constexpr uint32_t XLAYOUT = (C == 64) ? kLayoutSw128 : kLayoutSw64;
    static constexpr int IMG_MLP1 = SLICE * XROWB;         // B of mlp1: 32 rows x C
    static constexpr int IMG_TAP = SLICE * YROWB;          // B of one conv tap: 32 x 32 block diagonal
    static constexpr int IMG_MLP2 = C * YROWB;             // B of mlp2: C rows x 32
    static constexpr int IMG = IMG_MLP1 + 9 * IMG_TAP + IMG_MLP2;
    static constexpr int X_BYTES = ROWS * XROWB, Y_BYTES = ROWS * YROWB;
    static constexpr int SMEM = 1024 + X_BYTES + Y_BYTES + 2 * IMG;
    static constexpr int TMEM_P = 0, TMEM_Q = MT * SLICE;
    static constexpr int TMEM_USED = MT * SLICE + MT * C;
    static constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : (TMEM_USED <= 64 ? 64 : (TMEM_USED <= 128 ? 128 : (TMEM_USED <= 256 ? 256 : 512)));
    static_assert(SMEM <= 113 * 1024, "k_vit_ffn: two CTAs per SM");
    static_assert(TMEM_COLS <= 256, "k_vit_ffn: two CTAs per SM share the 512 TMEM columns");
};

// swizzled byte offset of 16-byte chunk `c16` of row `row` in a K-major tile whose base is 1024-byte aligned
template <int ROWB>
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t c16) {
    return row * ROWB + (((ROWB == 128) ? (c16 ^ (row & 7u)) : (c16 ^ ((row >> 1) & 3u))) << 4);
}

__device__ __forceinline__ float gelu_exact(float a) { return 0.5f * a * (1.f + erff(a * 0.70710678118654752440f)); }

// tokens bf16 [B][TH*TW][C] -> out bf16 [B][TH*TW][C]; w_img: [NSLICE][IMG] bytes (pre-swizzled, see tc.pack_vit_ffn);
// fbias fp32: [CE] mlp1 bias, [CE] conv bias, [C] mlp2 bias, [C] LayerNorm gamma, [C] beta
template <int C, int TH, int TW>
__global__ void __launch_bounds__(256, 2)
k_vit_ffn(const __nv_bfloat16* __restrict__ tokens, const uint8_t* __restrict__ w_img, const float* __restrict__ fbias,
          __nv_bfloat16* __restrict__ out, int B, float eps) {
    using Cfg = FfnCfg<C, TH, TW>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_x = smem;                                   // [ROWS][XROWB]
    uint8_t* s_y = smem + Cfg::X_BYTES;                    // [ROWS][YROWB]
    uint8_t* s_w = s_y + Cfg::Y_BYTES;                     // 2 x IMG
    static_assert(Cfg::X_BYTES % 1024 == 0 && Cfg::Y_BYTES % 512 == 0 && Cfg::IMG % 1024 == 0, "operand alignment");
    __shared__ __align__(8) uint64_t s_bar_w[2];
    __shared__ __align__(8) uint64_t s_bar_mma;
    __shared__ uint32_t s_tmem;
    __shared__ float s_b1[Cfg::CE], s_bdw[Cfg::CE], s_b2[C], s_gamma[C], s_beta[C];
    __shared__ float s_ln[128][2][2];

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int q4 = warp & 3, half = warp >> 2;            // TMEM lane quarter this warp may touch; which half of a tile's columns
    if (warp == 0) tmem_alloc(&s_tmem, Cfg::TMEM_COLS);
    if (t == 0) {
        mbar_init(&s_bar_w[0], 1);
        mbar_init(&s_bar_w[1], 1);
        mbar_init(&s_bar_mma, 1);
        fence_barrier_init();
    }
    for (int i = t; i < Cfg::CE; i += 256) {
        s_b1[i] = fbias[i];
        s_bdw[i] = fbias[Cfg::CE + i];
    }
    if (t < C) {
        s_b2[t] = fbias[2 * Cfg::CE + t];
        s_gamma[t] = fbias[2 * Cfg::CE + C + t];
        s_beta[t] = fbias[2 * Cfg::CE + 2 * C + t];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const uint32_t x_base = smem_u32(s_x), y_base = smem_u32(s_y), w_base = smem_u32(s_w);
    constexpr uint32_t idesc_s = make_idesc_bf16(128, Cfg::SLICE);     // N = 32: mlp1 slice, conv
    constexpr uint32_t idesc_c = make_idesc_bf16(128, C);              // N = C : mlp2
    uint32_t ph_mma = 0, ph_w0 = 0, ph_w1 = 0;
    // base descriptors: a descriptor's start-address field counts 16-byte units, so an operand `off` bytes further is
    // desc + (off >> 4) -- compile-time constants below; rebuilding descriptors per MMA costs more than a 32-wide MMA takes
    const uint64_t dx0 = make_smem_desc(x_base, 8 * Cfg::XROWB, Cfg::XLAYOUT);
    const uint64_t dy0 = make_smem_desc(y_base, 8 * Cfg::YROWB, kLayoutSw64);

    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        // ---- zero the padded grids, then drop the sample's tokens into X ------------------------------
        {
            uint4* z = reinterpret_cast<uint4*>(s_x);
            const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
            for (int i = t; i < (Cfg::X_BYTES + Cfg::Y_BYTES) / 16; i += 256) z[i] = zero;
        }
        __syncthreads();
        {
            constexpr int CPR = C / 8;                     // 16-byte chunks per token
            const uint4* src = reinterpret_cast<const uint4*>(tokens + (size_t)b * TH * TW * C);
            for (int i = t; i < TH * TW * CPR; i += 256) {
                const int tok = i / CPR, c16 = i - tok * CPR;
                const int yy = tok / TW, xx = tok - yy * TW;
                const uint32_t row = Cfg::LEAD + (yy + 1) * Cfg::WP + xx + 1;
                *reinterpret_cast<uint4*>(s_x + swz<Cfg::XROWB>(row, c16)) = __ldg(src + i);
            }
        }
        if (t == 0) {
            mbar_expect_tx(&s_bar_w[0], Cfg::IMG);
            bulk_g2s(s_w, w_img, Cfg::IMG, &s_bar_w[0]);
        }
        fence_proxy_async();
        __syncthreads();

        for (int sl = 0; sl < Cfg::NSLICE; ++sl) {
            const int buf = sl & 1;
            const uint32_t wb = w_base + buf * Cfg::IMG;
            const uint64_t dwx = make_smem_desc(wb, 8 * Cfg::XROWB, Cfg::XLAYOUT);                     // mlp1 weights (rows of C)
            const uint64_t dwy = make_smem_desc(wb + Cfg::IMG_MLP1, 8 * Cfg::YROWB, kLayoutSw64);      // conv taps, then mlp2 weights
            if (warp == 0) {                               // whole warp convergent, one elected lane issues
                if (sl + 1 < Cfg::NSLICE && elect_one()) { // the other buffer's last readers (slice sl-1) have completed
                    mbar_expect_tx(&s_bar_w[buf ^ 1], Cfg::IMG);
                    bulk_g2s(s_w + (buf ^ 1) * Cfg::IMG, w_img + (size_t)(sl + 1) * Cfg::IMG, Cfg::IMG, &s_bar_w[buf ^ 1]);
                }
                __syncwarp();
                mbar_wait(&s_bar_w[buf], buf ? ph_w1 : ph_w0);
                tc_fence_after();
                if (elect_one()) {
                    // mlp1 slice: P[mt] = X[mt] * W1_slice^T
#pragma unroll
                    for (int mt = 0; mt < Cfg::MT; ++mt)
#pragma unroll
                        for (int k = 0; k < C / 16; ++k)
                            umma_bf16(tmem_base + Cfg::TMEM_P + mt * Cfg::SLICE, dx0 + (uint64_t)(((Cfg::LEAD + Cfg::G0 + mt * 128) * Cfg::XROWB + k * 32) >> 4),
                                      dwx + (uint64_t)(k * 2), idesc_s, k != 0);
                    umma_commit(&s_bar_mma);
                }
                __syncwarp();
            }
            if (buf) ph_w1 ^= 1; else ph_w0 ^= 1;
            mbar_wait(&s_bar_mma, ph_mma);
            ph_mma ^= 1;
            tc_fence_after();
            // ---- epilogue 1: + bias, zero at the pad positions (the conv pads ITS input with zeros), bf16 -> Y
#pragma unroll 1
            for (int mt = 0; mt < Cfg::MT; ++mt) {
                uint32_t r[16];
                tmem_ld_32x16(tmem_base + ((uint32_t)(q4 * 32) << 16) + Cfg::TMEM_P + mt * Cfg::SLICE + half * 16, r);
                tmem_ld_wait();
                const int g = Cfg::G0 + mt * 128 + q4 * 32 + lane;
                const int gy = g / Cfg::WP, gx = g - gy * Cfg::WP;
                const bool valid = gx != 0 && gy >= 1 && gy <= TH;
                const float* bb = s_b1 + sl * Cfg::SLICE + half * 16;
                uint32_t pk[8];
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    pk[e] = valid ? pack_bf16x2(__uint_as_float(r[2 * e]) + bb[2 * e], __uint_as_float(r[2 * e + 1]) + bb[2 * e + 1]) : 0u;
                const uint32_t row = Cfg::LEAD + g;
                *reinterpret_cast<uint4*>(s_y + swz<Cfg::YROWB>(row, half * 2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(s_y + swz<Cfg::YROWB>(row, half * 2 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            // ---- grouped 3x3 conv of the slice: nine shifted GEMMs against block-diagonal taps
            if (warp == 0) {
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int mt = 0; mt < Cfg::MT; ++mt)
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const int shift = (tap / 3 - 1) * Cfg::WP + (tap % 3 - 1);
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                umma_bf16(tmem_base + Cfg::TMEM_P + mt * Cfg::SLICE,
                                          dy0 + (uint64_t)(((Cfg::LEAD + Cfg::G0 + mt * 128 + shift) * Cfg::YROWB + k * 32) >> 4),
                                          dwy + (uint64_t)((tap * Cfg::IMG_TAP + k * 32) >> 4), idesc_s, (tap | k) != 0);
                        }
                    umma_commit(&s_bar_mma);
                }
                __syncwarp();
            }
            mbar_wait(&s_bar_mma, ph_mma);
            ph_mma ^= 1;
            tc_fence_after();
            // ---- epilogue 2: + bias, exact GELU, bf16 -> Y in place (every conv MMA that read Y has completed)
#pragma unroll 1
            for (int mt = 0; mt < Cfg::MT; ++mt) {
                uint32_t r[16];
                tmem_ld_32x16(tmem_base + ((uint32_t)(q4 * 32) << 16) + Cfg::TMEM_P + mt * Cfg::SLICE + half * 16, r);
                tmem_ld_wait();
                const float* bb = s_bdw + sl * Cfg::SLICE + half * 16;
                uint32_t pk[8];
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    pk[e] = pack_bf16x2(gelu_exact(__uint_as_float(r[2 * e]) + bb[2 * e]), gelu_exact(__uint_as_float(r[2 * e + 1]) + bb[2 * e + 1]));
                const uint32_t row = Cfg::LEAD + Cfg::G0 + mt * 128 + q4 * 32 + lane;
                *reinterpret_cast<uint4*>(s_y + swz<Cfg::YROWB>(row, half * 2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(s_y + swz<Cfg::YROWB>(row, half * 2 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            // ---- mlp2 slice: Q[mt] += Y[mt] * W2_slice^T
            if (warp == 0) {
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int mt = 0; mt < Cfg::MT; ++mt)
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            umma_bf16(tmem_base + Cfg::TMEM_Q + mt * C, dy0 + (uint64_t)(((Cfg::LEAD + Cfg::G0 + mt * 128) * Cfg::YROWB + k * 32) >> 4),
                                      dwy + (uint64_t)((9 * Cfg::IMG_TAP + k * 32) >> 4), idesc_c, (sl | k) != 0);
                    umma_commit(&s_bar_mma);
                }
                __syncwarp();
            }
            mbar_wait(&s_bar_mma, ph_mma);          // Y and this weight buffer are free again
            ph_mma ^= 1;
            tc_fence_after();
        }
        // ---- final epilogue: + bias + residual -> LayerNorm(C) -> bf16 tokens; two threads share a row (C/2 columns each)
        constexpr int HC = C / 2;
#pragma unroll 1
        for (int mt = 0; mt < Cfg::MT; ++mt) {
            float v[HC];
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + Cfg::TMEM_Q + mt * C + half * HC;
            if (HC == 16) {
                uint32_t r[16];
                tmem_ld_32x16(taddr, r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
            } else {
                uint32_t r[32];
                tmem_ld_32x32(taddr, r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < HC; ++e) v[e] = __uint_as_float(r[e]);
            }
            const int rloc = q4 * 32 + lane;
            const int g = Cfg::G0 + mt * 128 + rloc;
            const int gy = g / Cfg::WP, gx = g - gy * Cfg::WP;
            const bool valid = gx != 0 && gy >= 1 && gy <= TH;
            const size_t tok = (size_t)b * TH * TW + (size_t)(gy - 1) * TW + (gx - 1);
            float sum = 0.f;
            if (valid) {
                const uint4* res = reinterpret_cast<const uint4*>(tokens + tok * C + half * HC);
#pragma unroll
                for (int c8 = 0; c8 < HC / 8; ++c8) {
                    const uint4 rr = __ldg(res + c8);
                    const __nv_bfloat162* pr = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __bfloat1622float2(pr[e]);
                        v[c8 * 8 + 2 * e] += f.x + s_b2[half * HC + c8 * 8 + 2 * e];
                        v[c8 * 8 + 2 * e + 1] += f.y + s_b2[half * HC + c8 * 8 + 2 * e + 1];
                    }
                }
#pragma unroll
                for (int e = 0; e < HC; ++e) sum += v[e];
            }
            s_ln[rloc][half][0] = sum;
            __syncthreads();
            const float mean = (s_ln[rloc][0][0] + s_ln[rloc][1][0]) / (float)C;
            float var = 0.f;
            if (valid) {
#pragma unroll
                for (int e = 0; e < HC; ++e) {
                    v[e] -= mean;
                    var = fmaf(v[e], v[e], var);
                }
            }
            s_ln[rloc][half][1] = var;
            __syncthreads();
            const float rstd = rsqrtf((s_ln[rloc][0][1] + s_ln[rloc][1][1]) / (float)C + eps);
            if (valid) {
                uint4* o = reinterpret_cast<uint4*>(out + tok * C + half * HC);
#pragma unroll
                for (int c8 = 0; c8 < HC / 8; ++c8) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = half * HC + c8 * 8 + 2 * e;
                        pk[e] = pack_bf16x2(v[c8 * 8 + 2 * e] * rstd * s_gamma[c] + s_beta[c], v[c8 * 8 + 2 * e + 1] * rstd * s_gamma[c + 1] + s_beta[c + 1]);
                    }
                    o[c8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
            __syncthreads();
        }
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// EfficientSelfAttention after the spatial reduction (learner/ViTsubmodules.py:74-83) + the residual of :144, one launch:
//     out = x + finalLayer( softmax(Q K^T / sqrt(d)) V ),   Q = query(x)
// K, V = keyValueExtractor(LayerNorm(cn1(x))) are 2..6 tokens per sample, computed by the launches in front
// (kv bf16 [B, n_kv, 2C], layout [kv][head][d] like ViTsubmodules.py:74). Per tile of 128 tokens: Q on tcgen05 ->
// TMEM -> every thread owns one token: few-key softmax in registers -> the attention row goes back to shared memory as
// the A operand of the final projection -> tcgen05 -> + bias + residual -> bf16. Round 1 ran this as three launches
// (two GEMMs and a CUDA-core attention kernel) with q and the attention output going through HBM/L2.
template <int C, int HEADS>
__global__ void __launch_bounds__(128)
k_vit_attn(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ kv, const uint8_t* __restrict__ w_img,
           const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, long long rows, int N, int n_kv, int n_tiles) {
    constexpr int ROWB = C * 2;
    constexpr int DH = C / HEADS;
    static_assert(DH == 32, "head dimension 32 (both stages of the reference's ViT)");
    constexpr uint32_t LAYOUT = (C == 64) ? kLayoutSw128 : kLayoutSw64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_a = smem;                         // [128][ROWB]  x tile
    uint8_t* s_p = s_a + 128 * ROWB;             // [128][ROWB]  attention rows
    uint8_t* s_wq = s_p + 128 * ROWB;            // [C][ROWB]
    uint8_t* s_wf = s_wq + C * ROWB;             // [C][ROWB]
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_bq[C], s_bf[C];
    const int t = threadIdx.x, warp = t >> 5;
    if (warp == 0) tmem_alloc(&s_tmem, 2 * C);
    if (t == 0) {
        mbar_init(&s_bar, 1);
        fence_barrier_init();
    }
    if (t < C) {
        s_bq[t] = bias[t];
        s_bf[t] = bias[C + t];
    }
    for (int i = t; i < 2 * C * ROWB / 16; i += 128) reinterpret_cast<uint4*>(s_wq)[i] = __ldg(reinterpret_cast<const uint4*>(w_img) + i);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const uint64_t da = make_smem_desc(smem_u32(s_a), 8 * ROWB, LAYOUT), dp = make_smem_desc(smem_u32(s_p), 8 * ROWB, LAYOUT);
    const uint64_t dq = make_smem_desc(smem_u32(s_wq), 8 * ROWB, LAYOUT), df = make_smem_desc(smem_u32(s_wf), 8 * ROWB, LAYOUT);
    constexpr uint32_t idesc = make_idesc_bf16(128, C);
    const float scale = rsqrtf((float)DH);
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row = (long long)tile * 128 + t;
        const bool live = row < rows;
        {   // x tile -> swizzled A operand
            const uint4* src = reinterpret_cast<const uint4*>(x + row * C);
#pragma unroll
            for (int c16 = 0; c16 < C / 8; ++c16)
                *reinterpret_cast<uint4*>(s_a + swz<ROWB>(t, c16)) = live ? __ldg(src + c16) : make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < C / 16; ++k) umma_bf16(tmem_base, da + (uint64_t)(k * 2), dq + (uint64_t)(k * 2), idesc, k != 0);
                umma_commit(&s_bar);
            }
            __syncwarp();
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- few-key attention of this thread's token
        {
            const long long b = live ? row / N : 0;
            const __nv_bfloat16* kvb = kv + b * (long long)n_kv * 2 * C;
#pragma unroll
            for (int h = 0; h < HEADS; ++h) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + h * DH, r);
                tmem_ld_wait();
                float q[DH];
#pragma unroll
                for (int d = 0; d < DH; ++d) q[d] = __uint_as_float(r[d]) + s_bq[h * DH + d];
                float sc[8];
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    sc[j] = -INFINITY;
                    if (j < n_kv) {
                        const uint4* kp = reinterpret_cast<const uint4*>(kvb + (long long)j * 2 * C + h * DH);
                        float dot = 0.f;
#pragma unroll
                        for (int c8 = 0; c8 < DH / 8; ++c8) {
                            const uint4 kk = __ldg(kp + c8);
                            const __nv_bfloat162* pk = reinterpret_cast<const __nv_bfloat162*>(&kk);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __bfloat1622float2(pk[e]);
                                dot = fmaf(q[c8 * 8 + 2 * e], f.x, dot);
                                dot = fmaf(q[c8 * 8 + 2 * e + 1], f.y, dot);
                            }
                        }
                        sc[j] = dot * scale;
                        mx = fmaxf(mx, sc[j]);
                    }
                }
                float den = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    sc[j] = j < n_kv ? __expf(sc[j] - mx) : 0.f;
                    den += sc[j];
                }
                const float inv = 1.f / den;
                float att[DH];
#pragma unroll
                for (int d = 0; d < DH; ++d) att[d] = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < n_kv) {
                        const uint4* vp = reinterpret_cast<const uint4*>(kvb + (long long)j * 2 * C + C + h * DH);
                        const float pj = sc[j] * inv;
#pragma unroll
                        for (int c8 = 0; c8 < DH / 8; ++c8) {
                            const uint4 vv = __ldg(vp + c8);
                            const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&vv);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __bfloat1622float2(pv[e]);
                                att[c8 * 8 + 2 * e] = fmaf(pj, f.x, att[c8 * 8 + 2 * e]);
                                att[c8 * 8 + 2 * e + 1] = fmaf(pj, f.y, att[c8 * 8 + 2 * e + 1]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int c8 = 0; c8 < DH / 8; ++c8)
                    *reinterpret_cast<uint4*>(s_p + swz<ROWB>(t, h * (DH / 8) + c8)) =
                        live ? make_uint4(pack_bf16x2(att[c8 * 8], att[c8 * 8 + 1]), pack_bf16x2(att[c8 * 8 + 2], att[c8 * 8 + 3]),
                                          pack_bf16x2(att[c8 * 8 + 4], att[c8 * 8 + 5]), pack_bf16x2(att[c8 * 8 + 6], att[c8 * 8 + 7]))
                             : make_uint4(0u, 0u, 0u, 0u);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < C / 16; ++k) umma_bf16(tmem_base + C, dp + (uint64_t)(k * 2), df + (uint64_t)(k * 2), idesc, k != 0);
                umma_commit(&s_bar);
            }
            __syncwarp();
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- final projection epilogue: + bias + residual -> bf16
#pragma unroll
        for (int c0 = 0; c0 < C; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + C + c0, r);
            tmem_ld_wait();
            if (live) {
                const uint4* res = reinterpret_cast<const uint4*>(x + row * C + c0);
                uint4* o = reinterpret_cast<uint4*>(out + row * C + c0);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    const uint4 rr = __ldg(res + c8);
                    const __nv_bfloat162* pr = reinterpret_cast<const __nv_bfloat162*>(&rr);
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __bfloat1622float2(pr[e]);
                        const int c = c0 + c8 * 8 + 2 * e;
                        pk[e] = pack_bf16x2(__uint_as_float(r[c8 * 8 + 2 * e]) + s_bf[c] + f.x, __uint_as_float(r[c8 * 8 + 2 * e + 1]) + s_bf[c + 1] + f.y);
                    }
                    o[c8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * C);
    }
}

template <int C, int HEADS>
static int launch_attn(const void* x, const void* kv, const void* w_img, const float* bias, void* out, long long rows, int N, int n_kv, cudaStream_t st) {
    constexpr int smem = 1024 + 2 * 128 * C * 2 + 2 * C * C * 2;
    EVFLY_SMEM_ATTR(smem, k_vit_attn<C, HEADS>);
    const long long tiles = (rows + 127) / 128;
    const long long cap = (long long)kNumSMs * (C == 32 ? 6 : 3);
    k_vit_attn<C, HEADS><<<(int)(tiles < cap ? tiles : cap), 128, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(kv),
                                                                            reinterpret_cast<const uint8_t*>(w_img), bias, reinterpret_cast<__nv_bfloat16*>(out), rows,
                                                                            N, n_kv, (int)tiles);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

template <int C, int TH, int TW>
static int launch_ffn(const void* tokens, const void* w_img, const float* fbias, void* out, int B, float eps, cudaStream_t st) {
    using Cfg = FfnCfg<C, TH, TW>;
    EVFLY_SMEM_ATTR(Cfg::SMEM, k_vit_ffn<C, TH, TW>);
    const int cap = 2 * kNumSMs;
    k_vit_ffn<C, TH, TW><<<B < cap ? B : cap, 256, Cfg::SMEM, st>>>(reinterpret_cast<const __nv_bfloat16*>(tokens), reinterpret_cast<const uint8_t*>(w_img),
                                                                  fbias, reinterpret_cast<__nv_bfloat16*>(out), B, eps);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

}  // namespace evfly

using namespace evfly;

extern "C" int64_t evfly_vit_ffn_image_bytes(int C) {
    if (C == 32) return (int64_t)FfnCfg<32, 15, 23>::NSLICE * FfnCfg<32, 15, 23>::IMG;
    if (C == 64) return (int64_t)FfnCfg<64, 8, 12>::NSLICE * FfnCfg<64, 8, 12>::IMG;
    return 0;
}

extern "C" int evfly_vit_ffn_bf16(const void* d_tokens, const void* d_w_img, const float* d_fbias, void* d_out, int B, int H, int W, int C,
                                  float eps, void* stream) {
    EVFLY_REQUIRE(d_tokens && d_w_img && d_fbias && d_out && B >= 0, "vit_ffn_bf16: bad argument");
    EVFLY_REQUIRE((reinterpret_cast<uintptr_t>(d_w_img) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_tokens) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0,
                  "vit_ffn_bf16: operands must be 16-byte aligned");
    if (B == 0) return EVFLY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 32 && H == 15 && W == 23) return launch_ffn<32, 15, 23>(d_tokens, d_w_img, d_fbias, d_out, B, eps, st);
    if (C == 64 && H == 8 && W == 12) return launch_ffn<64, 8, 12>(d_tokens, d_w_img, d_fbias, d_out, B, eps, st);
    set_error("vit_ffn_bf16: the fused MixFFN kernel is instantiated for the two stages of LSTMNetVIT / ViT (15x23x32 and 8x12x64), got %dx%dx%d", H, W, C);
    return EVFLY_ERR_UNSUPPORTED;
}

extern "C" int evfly_vit_attn_bf16(const void* d_x, const void* d_kv, const void* d_w_img, const float* d_bias, void* d_out, int64_t B, int N, int C,
                                   int heads, int n_kv, void* stream) {
    EVFLY_REQUIRE(d_x && d_kv && d_w_img && d_bias && d_out && B >= 0 && N > 0 && n_kv >= 1 && n_kv <= 8, "vit_attn_bf16: bad argument (1 <= n_kv <= 8)");
    EVFLY_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_kv) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_w_img) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(d_out) & 15) == 0, "vit_attn_bf16: operands must be 16-byte aligned");
    if (B == 0) return EVFLY_OK;
    const long long rows = (long long)B * N;
    EVFLY_REQUIRE(rows < (1ll << 31) * 64, "vit_attn_bf16: too many tokens");
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 32 && heads == 1) return launch_attn<32, 1>(d_x, d_kv, d_w_img, d_bias, d_out, rows, N, n_kv, st);
    if (C == 64 && heads == 2) return launch_attn<64, 2>(d_x, d_kv, d_w_img, d_bias, d_out, rows, N, n_kv, st);
    set_error("vit_attn_bf16: instantiated for (C, heads) = (32, 1) and (64, 2) (head dimension 32), got (%d, %d)", C, heads);
    return EVFLY_ERR_UNSUPPORTED;
}
