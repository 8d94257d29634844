// vit_bf16.cu -- bf16 path of the ViT / ViT-LSTM stages (learner/ViTsubmodules.py,
// learner/vitfly_models.py:132-150). Tokens are bf16 [B, N, C] = NHWC on the stage's H' x W' grid.
// The Linear layers (query, keyValueExtractor, finalLayer, mlp1, mlp2) run on evfly_tc_conv_bf16;
// this file holds what is not a GEMM: the (overlap-)patch-embedding convolution fused with its
// LayerNorm, LayerNorm, the few-key attention, the grouped 3x3 conv + GELU of MixFFN, and the LSTM
// scan with W_hh resident in shared memory.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace evfly {

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

// ---------------------------------------------------------------------------------------
// conv (k x k, stride s, pad p) + bias + LayerNorm(Cout) -> bf16 tokens [B, OH*OW, Cout]
// fp32 NCHW input with Cin == 1 (the depth image): one warp per output token; lane l owns channels
// l, l+32 (Cout <= 64). Weights fp32 [K][Cout], K = (kh*k + kw)*Cin + ci. The bf16 NHWC-input
// variant is k_patch_embed_ln_block below.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_patch_embed_ln(const float* __restrict__ xin, const float* __restrict__ w, const float* __restrict__ bias,
                 const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                 int B, int H, int W, int Cin, int Cout, int k, int s, int p, int OH, int OW, float eps) {
    const long long tok = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long total = (long long)B * OH * OW;
    if (tok >= total) return;
    const int lane = threadIdx.x & 31;
    const int ow = (int)(tok % OW), oh = (int)((tok / OW) % OH);
    const long long b = tok / ((long long)OW * OH);
    const bool two = Cout > 32;
    float a0 = bias[lane], a1 = two ? bias[lane + 32] : 0.f;
    for (int kh = 0; kh < k; ++kh) {
        const int ih = oh * s - p + kh;
        if ((unsigned)ih >= (unsigned)H) continue;
        for (int kw = 0; kw < k; ++kw) {
            const int iw = ow * s - p + kw;
            if ((unsigned)iw >= (unsigned)W) continue;
            const float* wp = w + (long long)((kh * k + kw) * Cin) * Cout;
            const float v = __ldg(xin + (b * H + ih) * (long long)W + iw);
            a0 = fmaf(v, wp[lane], a0);
            if (two) a1 = fmaf(v, wp[lane + 32], a1);
        }
    }
    // LayerNorm over Cout
    float sum = a0 + a1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const float mean = sum / (float)Cout;
    const float d0 = a0 - mean, d1 = two ? a1 - mean : 0.f;
    float var = d0 * d0 + d1 * d1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
    const float rstd = rsqrtf(var / (float)Cout + eps);
    __nv_bfloat16* o = out + tok * Cout;
    o[lane] = __float2bfloat16_rn(d0 * rstd * gamma[lane] + beta[lane]);
    if (two) o[lane + 32] = __float2bfloat16_rn(d1 * rstd * gamma[lane + 32] + beta[lane + 32]);
}

// bf16-NHWC-input variant with the K = k*k*Cin reduction split over the 8 warps of a CTA (one CTA
// per output token): the attention's reduction conv has K = 2048 and only 2..6 tokens per frame, so
// a warp per token would leave the GPU idle. The patch is staged in shared memory as fp32.
__global__ void __launch_bounds__(256)
k_patch_embed_ln_block(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                       int H, int W, int Cin, int Cout, int k, int s, int p, int OH, int OW, float eps) {
    extern __shared__ float s_patch[];            // [K] then [8][64] partial sums
    const int K = k * k * Cin;
    float* s_part = s_patch + K;
    const long long tok = blockIdx.x;
    const int ow = (int)(tok % OW), oh = (int)((tok / OW) % OH);
    const long long b = tok / ((long long)OW * OH);
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const int ci = i % Cin, t = i / Cin;
        const int ih = oh * s - p + t / k, iw = ow * s - p + t % k;
        float v = 0.f;
        if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) v = bf2f(x[((b * H + ih) * (long long)W + iw) * Cin + ci]);
        s_patch[i] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool two = Cout > 32;
    float a0 = 0.f, a1 = 0.f;
    const int per = (K + 7) / 8;
    const int k_end = min(K, (warp + 1) * per);
#pragma unroll 4
    for (int i = warp * per; i < k_end; ++i) {
        const float v = s_patch[i];
        a0 = fmaf(v, __ldg(w + (long long)i * Cout + lane), a0);
        if (two) a1 = fmaf(v, __ldg(w + (long long)i * Cout + lane + 32), a1);
    }
    s_part[warp * 64 + lane] = a0;
    s_part[warp * 64 + 32 + lane] = a1;
    __syncthreads();
    if (warp == 0) {
        a0 = bias[lane];
        a1 = two ? bias[lane + 32] : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            a0 += s_part[q * 64 + lane];
            a1 += s_part[q * 64 + 32 + lane];
        }
        float sum = a0 + a1;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        const float mean = sum / (float)Cout;
        const float d0 = a0 - mean, d1 = two ? a1 - mean : 0.f;
        float var = d0 * d0 + d1 * d1;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
        const float rstd = rsqrtf(var / (float)Cout + eps);
        __nv_bfloat16* o = out + tok * Cout;
        o[lane] = __float2bfloat16_rn(d0 * rstd * gamma[lane] + beta[lane]);
        if (two) o[lane + 32] = __float2bfloat16_rn(d1 * rstd * gamma[lane + 32] + beta[lane + 32]);
    }
}

// Few-token variant (batch-1 streaming: the reduction conv yields 2..6 tokens per frame with K up to 2048):
// a CLUSTER of 8 CTAs per token splits K eight ways, partial sums are reduced through distributed shared
// memory by the cluster's rank-0 CTA, which then applies bias + LayerNorm.
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256)
k_patch_embed_ln_cluster(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                         const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                         int H, int W, int Cin, int Cout, int k, int s, int p, int OH, int OW, float eps) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float s_patch[512];      // this CTA's K slice (<= 512 values)
    __shared__ float s_part[8 * 64];    // per-warp partial sums
    __shared__ float s_cta[64];         // this CTA's partial, read by rank 0 through DSMEM
    const int K = k * k * Cin;
    const int rank = (int)cluster.block_rank();
    const long long tok = blockIdx.x / 8;
    const int ow = (int)(tok % OW), oh = (int)((tok / OW) % OH);
    const long long b = tok / ((long long)OW * OH);
    const int per_cta = (K + 7) / 8;
    const int k0 = rank * per_cta, k1 = min(K, k0 + per_cta);
    for (int i = k0 + threadIdx.x; i < k1; i += blockDim.x) {
        const int ci = i % Cin, t = i / Cin;
        const int ih = oh * s - p + t / k, iw = ow * s - p + t % k;
        float v = 0.f;
        if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) v = bf2f(x[((b * H + ih) * (long long)W + iw) * Cin + ci]);
        s_patch[i - k0] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool two = Cout > 32;
    float a0 = 0.f, a1 = 0.f;
    const int per_warp = (k1 - k0 + 7) / 8;
    const int kw0 = k0 + warp * per_warp, kw1 = min(k1, kw0 + per_warp);
#pragma unroll 4
    for (int i = kw0; i < kw1; ++i) {
        const float v = s_patch[i - k0];
        a0 = fmaf(v, __ldg(w + (long long)i * Cout + lane), a0);
        if (two) a1 = fmaf(v, __ldg(w + (long long)i * Cout + lane + 32), a1);
    }
    s_part[warp * 64 + lane] = a0;
    s_part[warp * 64 + 32 + lane] = a1;
    __syncthreads();
    if (threadIdx.x < 64) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += s_part[q * 64 + threadIdx.x];
        s_cta[threadIdx.x] = t;
    }
    cluster.sync();
    if (rank == 0 && warp == 0) {
        a0 = bias[lane];
        a1 = two ? bias[lane + 32] : 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float* remote = cluster.map_shared_rank(s_cta, r);
            a0 += remote[lane];
            a1 += remote[lane + 32];
        }
        float sum = a0 + a1;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        const float mean = sum / (float)Cout;
        const float d0 = a0 - mean, d1 = two ? a1 - mean : 0.f;
        float var = d0 * d0 + d1 * d1;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
        const float rstd = rsqrtf(var / (float)Cout + eps);
        __nv_bfloat16* o = out + tok * Cout;
        o[lane] = __float2bfloat16_rn(d0 * rstd * gamma[lane] + beta[lane]);
        if (two) o[lane + 32] = __float2bfloat16_rn(d1 * rstd * gamma[lane + 32] + beta[lane + 32]);
    }
    cluster.sync();     // keep every CTA's shared memory alive until rank 0 has read it
}

// LayerNorm over C (32 or 64) of bf16 rows, one warp per row
__global__ void __launch_bounds__(256)
k_layernorm_bf16(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ y, long long rows, int C, float eps) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const bool two = C > 32;
    const float a0 = bf2f(x[row * C + lane]), a1 = two ? bf2f(x[row * C + lane + 32]) : 0.f;
    float sum = a0 + a1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const float mean = sum / (float)C;
    const float d0 = a0 - mean, d1 = two ? a1 - mean : 0.f;
    float var = d0 * d0 + d1 * d1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
    const float rstd = rsqrtf(var / (float)C + eps);
    y[row * C + lane] = __float2bfloat16_rn(d0 * rstd * gamma[lane] + beta[lane]);
    if (two) y[row * C + lane + 32] = __float2bfloat16_rn(d1 * rstd * gamma[lane + 32] + beta[lane + 32]);
}

// softmax(q k^T / sqrt(d)) v with n_kv <= 8 keys; one WARP per (token, head), lane = channel of the head
// (d = C/heads <= 32): coalesced 64-byte reads of q / k / v, dot products by warp shuffle.
__global__ void __launch_bounds__(256)
k_attention_small_bf16(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv, __nv_bfloat16* __restrict__ out,
                       long long B, int N, int C, int heads, int n_kv) {
    const long long i = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= B * N * heads) return;
    const int lane = threadIdx.x & 31;
    const int h = (int)(i % heads);
    const long long bn = i / heads, b = bn / N;
    const int d = C / heads;
    const bool on = lane < d;
    const float qv = on ? bf2f(q[bn * C + h * d + lane]) : 0.f;
    const __nv_bfloat16* kvb = kv + b * n_kv * 2 * C + h * d + lane;
    const float inv = rsqrtf((float)d);
    float sc[8], vv[8];
    float mx = -INFINITY;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        if (s < n_kv) {
            float dot = on ? qv * bf2f(kvb[(long long)s * 2 * C]) : 0.f;
            vv[s] = on ? bf2f(kvb[(long long)s * 2 * C + C]) : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            sc[s] = dot * inv;
            mx = fmaxf(mx, sc[s]);
        }
    }
    float den = 0.f, acc = 0.f;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        if (s < n_kv) {
            const float e = __expf(sc[s] - mx);
            den += e;
            acc = fmaf(e, vv[s], acc);
        }
    }
    if (on) out[bn * C + h * d + lane] = __float2bfloat16_rn(acc / den);
}

// MixFFN "depthwise" conv: groups = C, 8 -> 8 channels per group, 3x3, pad 1, + bias + exact GELU.
// x, y bf16 NHWC [B,H,W,8C]. grid.y = group; each thread one pixel of the group: 8 inputs per tap
// (one 16-byte load), 576 FMAs against the group's weights in shared memory ([tap][ci][co] fp32).
__global__ void __launch_bounds__(128)
k_dwconv3x3_gelu(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 __nv_bfloat16* __restrict__ y, long long B, int H, int W, int Ce) {
    __shared__ __align__(16) float s_w[9 * 8 * 8];
    __shared__ float s_b[8];
    const int g = blockIdx.y;
    // w is PyTorch's [Ce, 8, 3, 3]: row (g*8 + co), then ci, then tap
    for (int i = threadIdx.x; i < 576; i += blockDim.x) {
        const int co = i & 7, ci = (i >> 3) & 7, tap = i >> 6;
        s_w[i] = w[((long long)(g * 8 + co) * 8 + ci) * 9 + tap];
    }
    if (threadIdx.x < 8) s_b[threadIdx.x] = bias[g * 8 + threadIdx.x];
    __syncthreads();
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= B * H * W) return;
    const int pw = (int)(pix % W), ph = (int)((pix / W) % H);
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = s_b[c];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int ih = ph + tap / 3 - 1, iw = pw + tap % 3 - 1;
        if ((unsigned)ih >= (unsigned)H || (unsigned)iw >= (unsigned)W) continue;
        const uint4 raw = *reinterpret_cast<const uint4*>(x + (pix + (long long)(tap / 3 - 1) * W + (tap % 3 - 1)) * Ce + g * 8);
        const __nv_bfloat162* pr = reinterpret_cast<const __nv_bfloat162*>(&raw);
        float in[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 f = __bfloat1622float2(pr[q]);
            in[2 * q] = f.x;
            in[2 * q + 1] = f.y;
        }
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
            const float4 w0 = *reinterpret_cast<const float4*>(&s_w[(tap * 8 + ci) * 8]);
            const float4 w1 = *reinterpret_cast<const float4*>(&s_w[(tap * 8 + ci) * 8 + 4]);
            acc[0] = fmaf(in[ci], w0.x, acc[0]); acc[1] = fmaf(in[ci], w0.y, acc[1]);
            acc[2] = fmaf(in[ci], w0.z, acc[2]); acc[3] = fmaf(in[ci], w0.w, acc[3]);
            acc[4] = fmaf(in[ci], w1.x, acc[4]); acc[5] = fmaf(in[ci], w1.y, acc[5]);
            acc[6] = fmaf(in[ci], w1.z, acc[6]); acc[7] = fmaf(in[ci], w1.w, acc[7]);
        }
    }
    uint4 o;
    __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float a = acc[2 * q], b2 = acc[2 * q + 1];
        const float ga = 0.5f * a * (1.f + erff(a * 0.70710678118654752440f));
        const float gb = 0.5f * b2 * (1.f + erff(b2 * 0.70710678118654752440f));
        po[q] = __floats2bfloat162_rn(ga, gb);
    }
    *reinterpret_cast<uint4*>(y + pix * Ce + g * 8) = o;
}

// ---------------------------------------------------------------------------------------
// Tail of the ViT encoder (vitfly_models.py:136-142): cat([PixelShuffle(2)(s2), Upsample((2*H2, 2*W2), bilinear,
// align_corners=True)(s1)], dim=1) written as ONE dense NHWC bf16 tensor [B, 2*H2, 2*W2, ld] (channels
// [0, C2/4) shuffle, [C2/4, C2/4+C1) upsample, the rest zero) that the tensor-core 3x3 conv reads directly.
// One thread per (pixel, 8-channel chunk).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void bilin_src_align(int dst, int in, int out, int& i0, int& i1, float& l1) {
    const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    const float src = scale * (float)dst;
    i0 = (int)src;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

__global__ void __launch_bounds__(256)
k_shuffle_upsample_cat(const __nv_bfloat16* __restrict__ t2, int H2, int W2, int C2, const __nv_bfloat16* __restrict__ t1, int H1, int W1,
                       int C1, __nv_bfloat16* __restrict__ out, long long B, int ld) {
    const int OH = 2 * H2, OW = 2 * W2, chunks = ld >> 3, cs = C2 >> 2;
    const long long total = B * OH * OW * chunks;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ch = (int)(i % chunks);
    const long long pix = i / chunks;
    const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH);
    const long long b = pix / ((long long)OW * OH);
    const int c0 = ch * 8;
    float f[8];
    if (c0 < cs) {                     // PixelShuffle: out[c][2h+i][2w+j] = in[4c + 2i + j][h][w]
        const __nv_bfloat16* src = t2 + ((b * H2 + (oy >> 1)) * W2 + (ox >> 1)) * C2 + (oy & 1) * 2 + (ox & 1);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = c0 + e < cs ? __bfloat162float(src[(c0 + e) * 4]) : 0.f;
    } else if (c0 < cs + C1) {         // bilinear, align_corners=True
        int h0, h1, w0, w1;
        float lh, lw;
        bilin_src_align(oy, H1, OH, h0, h1, lh);
        bilin_src_align(ox, W1, OW, w0, w1, lw);
        const __nv_bfloat16* base = t1 + b * (long long)H1 * W1 * C1 + (c0 - cs);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (c0 - cs + e < C1) {
                const float v00 = __bfloat162float(base[((long long)h0 * W1 + w0) * C1 + e]), v01 = __bfloat162float(base[((long long)h0 * W1 + w1) * C1 + e]);
                const float v10 = __bfloat162float(base[((long long)h1 * W1 + w0) * C1 + e]), v11 = __bfloat162float(base[((long long)h1 * W1 + w1) * C1 + e]);
                f[e] = (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);
            } else {
                f[e] = 0.f;
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
    }
    uint4 o;
    __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) po[q] = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
    *reinterpret_cast<uint4*>(out + pix * ld + c0) = o;
}

// Same op, 4 horizontally adjacent pixels per thread: the 576 weights of the group are read from shared memory
// once per 4 pixels (LDS:FMA 1:16 instead of 1:4 -- the 1-pixel version is shared-memory-issue bound) and the
// 3 x 6 input window is loaded once (18 loads instead of 36).
__global__ void __launch_bounds__(128)
k_dwconv3x3_gelu_x4(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ y, long long B, int H, int W, int Ce) {
    __shared__ __align__(16) float s_w[9 * 8 * 8];
    __shared__ float s_b[8];
    const int g = blockIdx.y;
    for (int i = threadIdx.x; i < 576; i += blockDim.x) {
        const int co = i & 7, ci = (i >> 3) & 7, tap = i >> 6;
        s_w[i] = w[((long long)(g * 8 + co) * 8 + ci) * 9 + tap];
    }
    if (threadIdx.x < 8) s_b[threadIdx.x] = bias[g * 8 + threadIdx.x];
    __syncthreads();
    const int WG = (W + 3) >> 2;
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= B * H * WG) return;
    const int wg = (int)(item % WG), ph = (int)((item / WG) % H);
    const long long b = item / ((long long)WG * H);
    const int w0 = wg * 4;
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[p][c] = s_b[c];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int ih = ph + kh - 1;
        if ((unsigned)ih >= (unsigned)H) continue;
        const __nv_bfloat16* row = x + ((b * H + ih) * W) * Ce + g * 8;
        float in[6][8];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int iw = w0 - 1 + j;
            uint4 raw = make_uint4(0, 0, 0, 0);
            if ((unsigned)iw < (unsigned)W) raw = *reinterpret_cast<const uint4*>(row + (long long)iw * Ce);
            const __nv_bfloat162* pr = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 f = __bfloat1622float2(pr[q]);
                in[j][2 * q] = f.x;
                in[j][2 * q + 1] = f.y;
            }
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
                const float4 wa = *reinterpret_cast<const float4*>(&s_w[((kh * 3 + kw) * 8 + ci) * 8]);
                const float4 wb = *reinterpret_cast<const float4*>(&s_w[((kh * 3 + kw) * 8 + ci) * 8 + 4]);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float v = in[p + kw][ci];
                    acc[p][0] = fmaf(v, wa.x, acc[p][0]); acc[p][1] = fmaf(v, wa.y, acc[p][1]);
                    acc[p][2] = fmaf(v, wa.z, acc[p][2]); acc[p][3] = fmaf(v, wa.w, acc[p][3]);
                    acc[p][4] = fmaf(v, wb.x, acc[p][4]); acc[p][5] = fmaf(v, wb.y, acc[p][5]);
                    acc[p][6] = fmaf(v, wb.z, acc[p][6]); acc[p][7] = fmaf(v, wb.w, acc[p][7]);
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if (w0 + p >= W) break;
        uint4 o;
        __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float a = acc[p][2 * q], b2 = acc[p][2 * q + 1];
            const float ga = 0.5f * a * (1.f + erff(a * 0.70710678118654752440f));
            const float gb = 0.5f * b2 * (1.f + erff(b2 * 0.70710678118654752440f));
            po[q] = __floats2bfloat162_rn(ga, gb);
        }
        *reinterpret_cast<uint4*>(y + ((b * H + ph) * W + w0 + p) * Ce + g * 8) = o;
    }
}

// ---------------------------------------------------------------------------------------
// LSTM layer over an unbatched sequence with W_hh^T resident in shared memory as bf16 pairs:
// s_w[(j/2) * 4H + r] = {W_hh[r][j], W_hh[r][j+1]}. One persistent CTA of 4H threads; thread r owns
// gate row r. h and c stay fp32 in shared memory across all T steps.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_lstm_seq_smemw(const float* __restrict__ gx, const __nv_bfloat162* __restrict__ whh_pairs, const float* __restrict__ h0,
                 const float* __restrict__ c0, float* __restrict__ hs, float* __restrict__ hT, float* __restrict__ cT, int T,
                 int H, int n_seq) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    const int G = 4 * H;
    // sequence blockIdx.x of n_seq: gx [T, n_seq, 4H], hs [T, n_seq, H], states [n_seq, H] (time-major batch)
    const int sq = blockIdx.x;
    gx += (long long)sq * G;
    hs += (long long)sq * H;
    if (h0) h0 += (long long)sq * H;
    if (c0) c0 += (long long)sq * H;
    if (hT) hT += (long long)sq * H;
    if (cT) cT += (long long)sq * H;
    const long long gstep = (long long)n_seq * G, hstep = (long long)n_seq * H;
    __nv_bfloat162* s_w = reinterpret_cast<__nv_bfloat162*>(sm_raw);        // [H/2][G]
    float* s_h = reinterpret_cast<float*>(sm_raw + (size_t)(H / 2) * G * 4);  // [H]
    float* s_c = s_h + H;                                                     // [H]
    float* s_g = s_c + H;                                                     // [G]
    for (int i = threadIdx.x; i < (H / 2) * G; i += blockDim.x) s_w[i] = whh_pairs[i];
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        s_h[j] = h0 ? h0[j] : 0.f;
        s_c[j] = c0 ? c0[j] : 0.f;
    }
    __syncthreads();
    const int r = threadIdx.x;
    float gx_next = (r < G && T > 0) ? gx[r] : 0.f;
    for (int t = 0; t < T; ++t) {
        if (r < G) {
            float acc0 = gx_next, acc1 = 0.f;
            if (t + 1 < T) gx_next = gx[(long long)(t + 1) * gstep + r];   // hide the L2 latency behind this step's dot product
#pragma unroll 8
            for (int j2 = 0; j2 < H / 2; ++j2) {
                const float2 wv = __bfloat1622float2(s_w[j2 * G + r]);
                const float2 hv = *reinterpret_cast<const float2*>(&s_h[2 * j2]);
                acc0 = fmaf(hv.x, wv.x, acc0);
                acc1 = fmaf(hv.y, wv.y, acc1);
            }
            s_g[r] = acc0 + acc1;
        }
        __syncthreads();
        if (r < H) {
            const float ig = 1.f / (1.f + __expf(-s_g[r]));
            const float fg = 1.f / (1.f + __expf(-s_g[H + r]));
            const float gg = tanhf(s_g[2 * H + r]);
            const float og = 1.f / (1.f + __expf(-s_g[3 * H + r]));
            const float c = fg * s_c[r] + ig * gg;
            const float h = og * tanhf(c);
            s_c[r] = c;
            s_h[r] = h;
            hs[(long long)t * hstep + r] = h;
        }
        __syncthreads();
    }
    if (r < H) {
        if (hT) hT[r] = s_h[r];
        if (cT) cT[r] = s_c[r];
    }
}

}  // namespace evfly

namespace evfly {
int patch_embed_tc_dispatch(const void* x, int x_is_f32, const float* w, const float* bias, const float* gamma, const float* beta, void* out,
                            long long tokens, int H, int W, int Cin, int Cout, int k, int s, int p, int OH, int OW, float eps, cudaStream_t st);
}
using namespace evfly;

extern "C" int evfly_patch_embed_ln_bf16(const void* d_x, int x_is_f32_nchw, const float* d_w_kc, const float* d_bias,
                                         const float* d_gamma, const float* d_beta, void* d_tokens, int B, int H, int W, int Cin,
                                         int Cout, int k, int stride, int pad, float eps, void* stream) {
    EVFLY_REQUIRE(d_x && d_w_kc && d_bias && d_gamma && d_beta && d_tokens, "patch_embed_ln_bf16: null pointer");
    EVFLY_REQUIRE(B >= 0 && H > 0 && W > 0 && k > 0 && stride > 0 && pad >= 0 && (Cout == 32 || Cout == 64), "patch_embed_ln_bf16: bad shape (Cout must be 32 or 64)");
    EVFLY_REQUIRE(x_is_f32_nchw ? Cin == 1 : (Cin % 32 == 0), "patch_embed_ln_bf16: fp32 input needs Cin == 1, bf16 NHWC input Cin %% 32 == 0");
    if (B == 0) return EVFLY_OK;
    const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
    EVFLY_REQUIRE(OH > 0 && OW > 0, "patch_embed_ln_bf16: empty output");
    const long long total = (long long)B * OH * OW;
    if (total >= 1024) {     // batches: im2col rows built in shared memory, GEMM on tcgen05, LayerNorm in the epilogue (tc_patch_embed.cu)
        const int rc = patch_embed_tc_dispatch(d_x, x_is_f32_nchw, d_w_kc, d_bias, d_gamma, d_beta, d_tokens, total, H, W, Cin, Cout, k, stride, pad,
                                               OH, OW, eps, (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
    const unsigned grid = (unsigned)ceil_div(total, 8);
    if (x_is_f32_nchw)
        k_patch_embed_ln<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(d_x), d_w_kc, d_bias, d_gamma, d_beta, reinterpret_cast<__nv_bfloat16*>(d_tokens), B, H, W, Cin, Cout, k, stride, pad, OH, OW, eps);
    else {
        const size_t smem = ((size_t)k * k * Cin + 8 * 64) * sizeof(float);
        EVFLY_REQUIRE(smem <= 48 * 1024, "patch_embed_ln_bf16: patch of %d values does not fit shared memory", k * k * Cin);
        EVFLY_REQUIRE(total < (1ll << 28), "patch_embed_ln_bf16: too many tokens");
        if (total <= 64 && k * k * Cin >= 512 && k * k * Cin <= 4096) {
            // a handful of tokens with a long reduction: 8-CTA cluster per token (split K, DSMEM reduction)
            k_patch_embed_ln_cluster<<<(unsigned)total * 8, 256, 0, (cudaStream_t)stream>>>(
                reinterpret_cast<const __nv_bfloat16*>(d_x), d_w_kc, d_bias, d_gamma, d_beta, reinterpret_cast<__nv_bfloat16*>(d_tokens), H, W, Cin, Cout,
                k, stride, pad, OH, OW, eps);
            EVFLY_LAUNCHED();
            return EVFLY_OK;
        }
        k_patch_embed_ln_block<<<(unsigned)total, 256, smem, (cudaStream_t)stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(d_x), d_w_kc, d_bias, d_gamma, d_beta, reinterpret_cast<__nv_bfloat16*>(d_tokens), H, W, Cin, Cout,
            k, stride, pad, OH, OW, eps);
    }
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_layernorm_bf16(const void* d_x, const float* d_gamma, const float* d_beta, void* d_y, int64_t rows, int C,
                                    float eps, void* stream) {
    EVFLY_REQUIRE(d_x && d_gamma && d_beta && d_y && rows >= 0 && (C == 32 || C == 64), "layernorm_bf16: bad argument (C must be 32 or 64)");
    if (rows == 0) return EVFLY_OK;
    k_layernorm_bf16<<<(unsigned)ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(d_x), d_gamma, d_beta, reinterpret_cast<__nv_bfloat16*>(d_y), rows, C, eps);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_attention_small_bf16(const void* d_q, const void* d_kv, void* d_out, int64_t B, int N, int C, int heads,
                                          int n_kv, void* stream) {
    EVFLY_REQUIRE(d_q && d_kv && d_out && B >= 0 && N > 0 && heads > 0 && C % heads == 0 && C / heads <= 32 && n_kv > 0 && n_kv <= 8,
                  "attention_small_bf16: bad argument (head dim <= 32, n_kv <= 8)");
    if (B == 0) return EVFLY_OK;
    const long long total = B * N * heads;
    k_attention_small_bf16<<<(unsigned)ceil_div(total, 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(d_q), reinterpret_cast<const __nv_bfloat16*>(d_kv), reinterpret_cast<__nv_bfloat16*>(d_out), B, N, C, heads, n_kv);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_dwconv3x3_gelu_nhwc_bf16(const void* d_x, const float* d_w, const float* d_bias, void* d_y, int64_t B, int H,
                                              int W, int Ce, void* stream) {
    EVFLY_REQUIRE(d_x && d_w && d_bias && d_y && B >= 0 && H > 0 && W > 0 && Ce % 8 == 0, "dwconv3x3_gelu_nhwc_bf16: bad argument");
    if (B == 0) return EVFLY_OK;
    const long long pixels = B * H * W;
    if (pixels >= 4096) {      // throughput shape: 4 pixels per thread
        dim3 grid4((unsigned)ceil_div(B * H * ((W + 3) / 4), 128), (unsigned)(Ce / 8));
        k_dwconv3x3_gelu_x4<<<grid4, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(d_x), d_w, d_bias,
                                                                    reinterpret_cast<__nv_bfloat16*>(d_y), B, H, W, Ce);
        EVFLY_LAUNCHED();
        return EVFLY_OK;
    }
    dim3 grid((unsigned)ceil_div(pixels, 128), (unsigned)(Ce / 8));
    k_dwconv3x3_gelu<<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(d_x), d_w, d_bias,
                                                             reinterpret_cast<__nv_bfloat16*>(d_y), B, H, W, Ce);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_lstm_seq_smemw(const float* d_gx, const void* d_whh_pairs, const float* d_h0, const float* d_c0, float* d_hs,
                                    float* d_hT, float* d_cT, int T, int H, int n_seq, void* stream) {
    EVFLY_REQUIRE(d_gx && d_whh_pairs && d_hs && T >= 0 && H > 0 && H % 2 == 0 && 4 * H <= 1024 && n_seq > 0, "lstm_seq_smemw: bad argument (H even, 4H <= 1024)");
    const size_t smem = (size_t)(H / 2) * 4 * H * 4 + (size_t)6 * H * 4;
    EVFLY_REQUIRE(smem <= 220 * 1024, "lstm_seq_smemw: W_hh does not fit shared memory (H=%d)", H);
    EVFLY_SMEM_ATTR(220 * 1024, k_lstm_seq_smemw);
    const int threads = ((4 * H + 31) / 32) * 32;
    k_lstm_seq_smemw<<<n_seq, threads, smem, (cudaStream_t)stream>>>(d_gx, reinterpret_cast<const __nv_bfloat162*>(d_whh_pairs), d_h0, d_c0,
                                                                     d_hs, d_hT, d_cT, T, H, n_seq);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_shuffle_upsample_cat_bf16(const void* d_t2, int H2, int W2, int C2, const void* d_t1, int H1, int W1, int C1, void* d_out,
                                               int64_t B, int ld, void* stream) {
    EVFLY_REQUIRE(d_t2 && d_t1 && d_out && B >= 0 && H2 > 0 && W2 > 0 && H1 > 0 && W1 > 0 && C2 % 32 == 0 && C1 % 8 == 0 && ld % 8 == 0 && ld >= C2 / 4 + C1,
                  "shuffle_upsample_cat_bf16: bad argument (C2 % 32 == 0, C1 % 8 == 0, ld >= C2/4 + C1)");
    if (B == 0) return EVFLY_OK;
    const long long total = B * 4 * H2 * W2 * (ld / 8);
    k_shuffle_upsample_cat<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(d_t2), H2, W2, C2, reinterpret_cast<const __nv_bfloat16*>(d_t1), H1, W1, C1,
        reinterpret_cast<__nv_bfloat16*>(d_out), B, ld);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}
