// nhwc_bf16.cu -- the memory-bound glue of the bf16 path: stem conv (Cin 1/2), 2x2 max-pool,
// bilinear skip resize, crop, ConvLSTM cell update and the NCHW<->pitch-grid converters.
// All of these are HBM/L2-streaming kernels with 16-byte (8 x bf16) accesses per thread.
#include "common.cuh"
#include <cuda_bf16.h>

namespace evfly {

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(t);
}

// ---- stem: fp32 NCHW (Cin 1|2) -> bf16 NHWC 32 channels ----------------------------------------
// one thread per output pixel; the 32 x Cin x 9 weights live in shared memory
__global__ void __launch_bounds__(256)
k_stem_conv3x3(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               uint4* __restrict__ out, int N, int Cin, int H, int W) {
    __shared__ float s_w[2 * 9 * 32];  // [ci][tap][co]
    __shared__ float s_b[32];
    for (int i = threadIdx.x; i < Cin * 9 * 32; i += blockDim.x) {
        const int co = i % 32, t = (i / 32) % 9, ci = i / (32 * 9);
        s_w[i] = w[(co * Cin + ci) * 9 + t];
    }
    if (threadIdx.x < 32) s_b[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const long long total = (long long)N * H * W;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int ow = (int)(i % W), oh = (int)((i / W) % H);
        const long long n = i / ((long long)W * H);
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = s_b[c];
        if (oh < H - 2 && ow < W - 2) {
            for (int ci = 0; ci < Cin; ++ci) {
                const float* xp = x + ((n * Cin + ci) * H + oh) * (long long)W + ow;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const float v = __ldg(xp + (t / 3) * W + (t % 3));
                    const float* wp = s_w + (ci * 9 + t) * 32;
#pragma unroll
                    for (int c = 0; c < 32; ++c) acc[c] = fmaf(v, wp[c], acc[c]);
                }
            }
        }
        uint4* o = out + i * 4;  // 32 bf16 = 64 bytes
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 pk;
            pk.x = pack_bf16x2(fmaxf(acc[q * 8 + 0], 0.f), fmaxf(acc[q * 8 + 1], 0.f));
            pk.y = pack_bf16x2(fmaxf(acc[q * 8 + 2], 0.f), fmaxf(acc[q * 8 + 3], 0.f));
            pk.z = pack_bf16x2(fmaxf(acc[q * 8 + 4], 0.f), fmaxf(acc[q * 8 + 5], 0.f));
            pk.w = pack_bf16x2(fmaxf(acc[q * 8 + 6], 0.f), fmaxf(acc[q * 8 + 7], 0.f));
            o[q] = pk;
        }
    }
}

__device__ __forceinline__ uint4 max_bf16x8(uint4 a, uint4 b) {
    uint4 r;
    __nv_bfloat162* pa = reinterpret_cast<__nv_bfloat162*>(&a);
    __nv_bfloat162* pb = reinterpret_cast<__nv_bfloat162*>(&b);
    __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
    return r;
}

__global__ void __launch_bounds__(256)
k_maxpool2x2_nhwc(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int Hp, int Wp, int OH, int OW, int C8) {
    const long long total = (long long)N * OH * OW * C8;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % C8);
        const long long pix = i / C8;
        const int ow = (int)(pix % OW), oh = (int)((pix / OW) % OH);
        const long long n = pix / ((long long)OW * OH);
        const uint4* p = x + ((n * Hp + 2 * oh) * (long long)Wp + 2 * ow) * C8 + c;
        const uint4 a = p[0], b = p[C8], d = p[(long long)Wp * C8], e = p[(long long)Wp * C8 + C8];
        y[i] = max_bf16x8(max_bf16x8(a, b), max_bf16x8(d, e));
    }
}

__device__ __forceinline__ void bilin_src(int dst, int in, int out, int& i0, int& i1, float& l1) {
    const float scale = (float)in / (float)out;  // align_corners = False
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

template <typename Idx>
__global__ void __launch_bounds__(256)
k_resize_bilinear_nhwc(const uint4* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int Hp, int Wp, int vh, int vw,
                       int C8, int OH, int OW, long long out_ld, int out_c0) {
    // Idx = unsigned when the element count fits 32 bits: 64-bit integer division costs ~100 instructions
    const Idx total = (Idx)N * (Idx)OH * (Idx)OW * (Idx)C8;
    const Idx stride = (Idx)gridDim.x * blockDim.x;
    for (Idx i = (Idx)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const Idx pix = i / (Idx)C8;
        const int c = (int)(i - pix * (Idx)C8);
        const Idx prow = pix / (Idx)OW;
        const int ow = (int)(pix - prow * (Idx)OW);
        const Idx n = prow / (Idx)OH;
        const int oh = (int)(prow - n * (Idx)OH);
        int h0, h1, w0, w1;
        float lh, lw;
        bilin_src(oh, vh, OH, h0, h1, lh);
        bilin_src(ow, vw, OW, w0, w1, lw);
        const uint4* base = x + (long long)n * Hp * Wp * C8 + c;
        const uint4 v00 = base[((long long)h0 * Wp + w0) * C8], v01 = base[((long long)h0 * Wp + w1) * C8];
        const uint4 v10 = base[((long long)h1 * Wp + w0) * C8], v11 = base[((long long)h1 * Wp + w1) * C8];
        const uint32_t* a = &v00.x; const uint32_t* b = &v01.x; const uint32_t* d = &v10.x; const uint32_t* e = &v11.x;
        uint4 o;
        uint32_t* po = &o.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 fa = unpack_bf16x2(a[q]), fb = unpack_bf16x2(b[q]), fd = unpack_bf16x2(d[q]), fe = unpack_bf16x2(e[q]);
            const float r0 = (1.f - lh) * ((1.f - lw) * fa.x + lw * fb.x) + lh * ((1.f - lw) * fd.x + lw * fe.x);
            const float r1 = (1.f - lh) * ((1.f - lw) * fa.y + lw * fb.y) + lh * ((1.f - lw) * fd.y + lw * fe.y);
            po[q] = pack_bf16x2(r0, r1);
        }
        *reinterpret_cast<uint4*>(y + (long long)pix * out_ld + out_c0 + c * 8) = o;
    }
}

__global__ void __launch_bounds__(256)
k_crop_nhwc(const uint4* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int Hp, int Wp, int C8, int h0, int w0, int OH,
            int OW, long long out_ld, int out_c0) {
    const long long total = (long long)N * OH * OW * C8;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % C8);
        const long long pix = i / C8;
        const int ow = (int)(pix % OW), oh = (int)((pix / OW) % OH);
        const long long n = pix / ((long long)OW * OH);
        const uint4 v = x[((n * Hp + h0 + oh) * (long long)Wp + w0 + ow) * C8 + c];
        *reinterpret_cast<uint4*>(y + pix * out_ld + out_c0 + c * 8) = v;
    }
}

__global__ void __launch_bounds__(256)
k_convlstm_pointwise_nhwc(const float* __restrict__ gates, float* __restrict__ c, __nv_bfloat16* __restrict__ h, long long P, int Ch) {
    const long long total = P * Ch;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long pix = i / Ch;
        const int ch = (int)(i - pix * Ch);
        const float* g = gates + pix * 4 * Ch;
        const float gi = g[ch], gf = g[Ch + ch], go = g[2 * Ch + ch], gg = g[3 * Ch + ch];
        const float iv = 1.f / (1.f + __expf(-gi)), fv = 1.f / (1.f + __expf(-gf)), ov = 1.f / (1.f + __expf(-go));
        const float cn = fv * c[i] + iv * tanhf(gg);
        c[i] = cn;
        h[i] = __float2bfloat16_rn(ov * tanhf(cn));
    }
}

__global__ void __launch_bounds__(256)
k_nchw_to_nhwc_bf16(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int C, int vh, int vw, int Hp, int Wp) {
    const long long total = (long long)N * Hp * Wp * C;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const int w = (int)(pix % Wp), h = (int)((pix / Wp) % Hp);
        const long long n = pix / ((long long)Wp * Hp);
        float v = 0.f;
        if (h < vh && w < vw) v = x[((n * C + c) * vh + h) * (long long)vw + w];
        y[i] = __float2bfloat16_rn(v);
    }
}

__global__ void __launch_bounds__(256)
k_nhwc_to_nchw_f32(const void* __restrict__ x, int src_is_f32, float* __restrict__ y, int N, int C, int vh, int vw, int Hp, int Wp) {
    const long long total = (long long)N * C * vh * vw;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int w = (int)(i % vw), h = (int)((i / vw) % vh);
        const int c = (int)((i / ((long long)vw * vh)) % C);
        const long long n = i / ((long long)vw * vh * C);
        const long long src = ((n * Hp + h) * (long long)Wp + w) * C + c;
        y[i] = src_is_f32 ? reinterpret_cast<const float*>(x)[src] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[src]);
    }
}

}  // namespace evfly

using namespace evfly;
static inline int g_ew(long long n) { return stream_grid(n, 256 * 2, 16); }

extern "C" int evfly_stem_conv3x3_fma_bf16(const float* d_x, const float* d_w, const float* d_bias, void* d_out, int N, int Cin,
                                       int H, int W, void* stream) {
    EVFLY_REQUIRE(d_x && d_w && d_bias && d_out && N >= 0 && (Cin == 1 || Cin == 2) && H >= 3 && W >= 3, "stem_conv3x3_fma_bf16: bad argument (Cin must be 1 or 2)");
    if (N == 0) return EVFLY_OK;
    k_stem_conv3x3<<<stream_grid((long long)N * H * W, 256, 16), 256, 0, (cudaStream_t)stream>>>(d_x, d_w, d_bias, reinterpret_cast<uint4*>(d_out), N, Cin, H, W);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_maxpool2x2_nhwc_bf16(const void* d_x, void* d_y, int N, int Hp, int Wp, int vh, int vw, int C, void* stream) {
    EVFLY_REQUIRE(d_x && d_y && N >= 0 && C % 8 == 0 && vh >= 2 && vw >= 2 && vh <= Hp && vw <= Wp, "maxpool2x2_nhwc_bf16: bad argument");
    if (N == 0) return EVFLY_OK;
    const int OH = vh / 2, OW = vw / 2;
    k_maxpool2x2_nhwc<<<g_ew((long long)N * OH * OW * (C / 8)), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(d_x), reinterpret_cast<uint4*>(d_y), N, Hp, Wp, OH, OW, C / 8);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_resize_bilinear_nhwc_bf16(const void* d_x, void* d_y, int N, int Hp, int Wp, int vh, int vw, int C, int OH,
                                               int OW, int64_t out_ld, int out_c0, void* stream) {
    EVFLY_REQUIRE(d_x && d_y && N >= 0 && C % 8 == 0 && vh > 0 && vw > 0 && vh <= Hp && vw <= Wp && OH > 0 && OW > 0 && out_ld % 8 == 0 && out_c0 % 8 == 0,
                  "resize_bilinear_nhwc_bf16: bad argument");
    if (N == 0) return EVFLY_OK;
    const long long total = (long long)N * OH * OW * (C / 8);
    if (total < (1ll << 31))
        k_resize_bilinear_nhwc<unsigned><<<g_ew(total), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const uint4*>(d_x), reinterpret_cast<__nv_bfloat16*>(d_y), N, Hp, Wp, vh, vw, C / 8, OH, OW, out_ld, out_c0);
    else
        k_resize_bilinear_nhwc<long long><<<g_ew(total), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const uint4*>(d_x), reinterpret_cast<__nv_bfloat16*>(d_y), N, Hp, Wp, vh, vw, C / 8, OH, OW, out_ld, out_c0);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_crop_nhwc_bf16(const void* d_x, void* d_y, int N, int Hp, int Wp, int C, int h0, int w0, int OH, int OW,
                                    int64_t out_ld, int out_c0, void* stream) {
    EVFLY_REQUIRE(d_x && d_y && N >= 0 && C % 8 == 0 && h0 >= 0 && w0 >= 0 && h0 + OH <= Hp && w0 + OW <= Wp && out_ld % 8 == 0 && out_c0 % 8 == 0,
                  "crop_nhwc_bf16: bad argument");
    if (N == 0) return EVFLY_OK;
    k_crop_nhwc<<<g_ew((long long)N * OH * OW * (C / 8)), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(d_x), reinterpret_cast<__nv_bfloat16*>(d_y), N, Hp, Wp, C / 8, h0, w0, OH, OW, out_ld, out_c0);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_convlstm_pointwise_nhwc(const float* d_gates, float* d_c, void* d_h_bf16, int64_t P, int Ch, void* stream) {
    EVFLY_REQUIRE(d_gates && d_c && d_h_bf16 && P > 0 && Ch > 0, "convlstm_pointwise_nhwc: bad argument");
    k_convlstm_pointwise_nhwc<<<g_ew(P * Ch), 256, 0, (cudaStream_t)stream>>>(d_gates, d_c, reinterpret_cast<__nv_bfloat16*>(d_h_bf16), P, Ch);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_nchw_f32_to_nhwc_bf16(const float* d_x, void* d_y, int N, int C, int vh, int vw, int Hp, int Wp, void* stream) {
    EVFLY_REQUIRE(d_x && d_y && N >= 0 && C > 0 && vh > 0 && vw > 0 && vh <= Hp && vw <= Wp, "nchw_f32_to_nhwc_bf16: bad argument");
    if (N == 0) return EVFLY_OK;
    k_nchw_to_nhwc_bf16<<<g_ew((long long)N * Hp * Wp * C), 256, 0, (cudaStream_t)stream>>>(d_x, reinterpret_cast<__nv_bfloat16*>(d_y), N, C, vh, vw, Hp, Wp);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_nhwc_to_nchw_f32(const void* d_x, int src_is_f32, float* d_y, int N, int C, int vh, int vw, int Hp, int Wp, void* stream) {
    EVFLY_REQUIRE(d_x && d_y && N >= 0 && C > 0 && vh > 0 && vw > 0 && vh <= Hp && vw <= Wp, "nhwc_to_nchw_f32: bad argument");
    if (N == 0) return EVFLY_OK;
    k_nhwc_to_nchw_f32<<<g_ew((long long)N * C * vh * vw), 256, 0, (cudaStream_t)stream>>>(d_x, src_is_f32, d_y, N, C, vh, vw, Hp, Wp);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}
