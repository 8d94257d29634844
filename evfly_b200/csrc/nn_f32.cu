// nn_f32.cu -- L3, the exact (fp32, CUDA-core) operator path of the model forward.
//
// This is the path whose outputs must match the reference's PyTorch-CPU forward within
// rtol 1e-5 (BASELINE.json north_star); the bf16 tensor-core path (nn_bf16.cu) is the fast one.
// Everything is a strided implicit-GEMM convolution plus a few small kernels, because the
// reference's layers are exactly that (learner/learner_models.py:373-414, ViTsubmodules.py,
// vitfly_models.py). Strides let flatten/transpose/permute/cat be expressed as addressing.
#include "common.cuh"
#include <math.h>

namespace evfly {

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case EVFLY_ACT_RELU: return v > 0.f ? v : 0.f;
        case EVFLY_ACT_LEAKY: return v > 0.f ? v : 0.01f * v;
        case EVFLY_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        case EVFLY_ACT_TANH: return tanhf(v);
        case EVFLY_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

// ---------------------------------------------------------------------------------------
// conv2d: implicit GEMM, 64 pixels x 64 output channels per CTA, K chunks of 16,
// 256 threads each owning a 4x4 register tile.
// ---------------------------------------------------------------------------------------
constexpr int CBM = 64, CBN = 64, CBK = 16;

__global__ void __launch_bounds__(256)
k_conv2d_f32(const evfly_conv2d_args a, const int OH, const int OW, const long long M, const int K,
             const int cin_g, const int cout_g) {
    __shared__ __align__(16) float As[CBK][CBM + 4];
    __shared__ __align__(16) float Bs[CBK][CBN + 4];
    __shared__ long long s_pix[CBM];
    __shared__ int s_ih0[CBM], s_iw0[CBM];
    __shared__ long long s_koff[CBK];
    __shared__ int s_kh[CBK], s_kw[CBK];

    const int tid = threadIdx.x;
    const int g = blockIdx.z;
    const long long m0 = (long long)blockIdx.x * CBM;
    const int n0 = blockIdx.y * CBN;
    const int KHW = a.KH * a.KW;

    if (tid < CBM) {
        const long long m = m0 + tid;
        if (m < M) {
            const int ow = (int)(m % OW);
            const int oh = (int)((m / OW) % OH);
            const long long n = m / ((long long)OW * OH);
            const int ih0 = oh * a.stride - a.pad, iw0 = ow * a.stride - a.pad;
            s_ih0[tid] = ih0;
            s_iw0[tid] = iw0;
            s_pix[tid] = n * a.xs[0] + (long long)ih0 * a.xs[2] + (long long)iw0 * a.xs[3];
        } else {
            s_ih0[tid] = -(1 << 28);  // fails every bounds test
            s_iw0[tid] = -(1 << 28);
            s_pix[tid] = 0;
        }
    }
    const bool chan_fast = (a.xs[1] == 1);
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const float* __restrict__ wg = a.w + (size_t)g * cout_g * K;

    for (int k0 = 0; k0 < K; k0 += CBK) {
        if (tid < CBK) {
            const int k = k0 + tid;
            if (k < K) {
                const int ci = k / KHW, r = k - ci * KHW;
                const int kh = r / a.KW, kw = r - kh * a.KW;
                s_kh[tid] = kh;
                s_kw[tid] = kw;
                s_koff[tid] = (long long)(g * cin_g + ci) * a.xs[1] + (long long)kh * a.xs[2] + (long long)kw * a.xs[3];
            } else {
                s_kh[tid] = -(1 << 28);
                s_kw[tid] = 0;
                s_koff[tid] = 0;
            }
        }
        __syncthreads();  // also: everyone is done computing on the previous As/Bs
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int kk, mm;
            if (chan_fast) { kk = tid & 15; mm = (tid >> 4) + 16 * j; }
            else           { mm = tid & 63; kk = (tid >> 6) + 4 * j; }
            const int ih = s_ih0[mm] + s_kh[kk], iw = s_iw0[mm] + s_kw[kk];
            float v = 0.f;
            if ((unsigned)ih < (unsigned)a.H && (unsigned)iw < (unsigned)a.W)
                v = __ldg(a.x + s_pix[mm] + s_koff[kk]);
            As[kk][mm] = v;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kk = tid & 15, nn = (tid >> 4) + 16 * j;
            const int co = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (co < cout_g && k < K) ? __ldg(wg + (size_t)co * K + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < CBK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w};
            const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
        const int ow = (int)(m % OW);
        const int oh = (int)((m / OW) % OH);
        const long long n = m / ((long long)OW * OH);
        const long long obase = n * a.ys[0] + (long long)oh * a.ys[2] + (long long)ow * a.ys[3];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = n0 + tx * 4 + j;
            if (co >= cout_g) continue;
            const int c = g * cout_g + co;
            float v = acc[i][j];
            if (a.bias) v += a.bias[c];
            v = apply_act(v, a.act);
            if (a.post_scale) v = v * a.post_scale[c] + a.post_shift[c];
            const long long o = obase + (long long)c * a.ys[1];
            if (a.res) v += a.res[o];
            a.y[o] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Tiny convolutions (batch-1: the 48->12 down_sample conv, heads): the 64x64-tile kernel would run on a
// handful of CTAs with a long serial K loop; here every output element gets its own thread.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_conv2d_small(const evfly_conv2d_args a, const int OH, const int OW, const long long M, const int K, const int cin_g, const int cout_g) {
    // one WARP per output element; the lanes split K and reduce by shuffle
    const long long idx = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (idx >= M * a.Cout) return;
    const int lane = threadIdx.x & 31;
    const long long m = idx % M;
    const int c = (int)(idx / M);
    const int g = c / cout_g;
    const int ow = (int)(m % OW), oh = (int)((m / OW) % OH);
    const long long n = m / ((long long)OW * OH);
    const int ih0 = oh * a.stride - a.pad, iw0 = ow * a.stride - a.pad;
    const float* __restrict__ wp = a.w + (long long)c * K;
    const float* __restrict__ xp = a.x + n * a.xs[0] + (long long)(g * cin_g) * a.xs[1];
    const int KHW = a.KH * a.KW;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) {
        const int ci = k / KHW, r = k - ci * KHW;
        const int kh = r / a.KW, kw = r - kh * a.KW;
        const int ih = ih0 + kh, iw = iw0 + kw;
        if ((unsigned)ih < (unsigned)a.H && (unsigned)iw < (unsigned)a.W)
            acc = fmaf(__ldg(xp + ci * a.xs[1] + ih * a.xs[2] + iw * a.xs[3]), __ldg(wp + k), acc);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane != 0) return;
    if (a.bias) acc += a.bias[c];
    acc = apply_act(acc, a.act);
    if (a.post_scale) acc = acc * a.post_scale[c] + a.post_shift[c];
    const long long o = n * a.ys[0] + (long long)c * a.ys[1] + (long long)oh * a.ys[2] + (long long)ow * a.ys[3];
    if (a.res) acc += a.res[o];
    a.y[o] = acc;
}

// ---------------------------------------------------------------------------------------
// Linear with a handful of rows (batch-1 streaming: the 4608->512 decoder, the LSTM input
// projection, the heads): weight-bandwidth bound, so one warp per output feature streams its
// weight row once (coalesced) and serves all M <= 8 input rows from shared memory.
// ---------------------------------------------------------------------------------------
constexpr int kSmallM = 8;
__global__ void __launch_bounds__(256)
k_linear_smallm(const float* __restrict__ x, long long x_ld, const float* __restrict__ w, const float* __restrict__ bias,
                const float* __restrict__ res, long long res_ld, float* __restrict__ y, long long y_ld, int M, int N, int K, int act) {
    extern __shared__ float s_x[];  // [M][K]
    for (int i = threadIdx.x; i < M * K; i += blockDim.x) s_x[i] = x[(long long)(i / K) * x_ld + (i % K)];
    __syncthreads();
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    const int lane = threadIdx.x & 31;
    float acc[kSmallM];
#pragma unroll
    for (int m = 0; m < kSmallM; ++m) acc[m] = 0.f;
    const float* wr = w + (long long)n * K;
    for (int k = lane; k < K; k += 32) {
        const float wv = __ldg(wr + k);
#pragma unroll
        for (int m = 0; m < kSmallM; ++m)
            if (m < M) acc[m] = fmaf(s_x[m * K + k], wv, acc[m]);
    }
#pragma unroll
    for (int m = 0; m < kSmallM; ++m) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], d);
    }
    if (lane == 0) {
        for (int m = 0; m < M; ++m) {
            float v = acc[m] + (bias ? bias[n] : 0.f);
            v = apply_act(v, act);
            if (res) v += res[(long long)m * res_ld + n];
            y[(long long)m * y_ld + n] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pool2d(const float* __restrict__ x, float* __restrict__ y, long long planes, int H, int W, int OH,
         int OW, int k, int s, int mode, int neg_in, int neg_out) {
    const long long total = planes * OH * OW;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int ow = (int)(i % OW), oh = (int)((i / OW) % OH);
        const long long p = i / ((long long)OW * OH);
        const float* src = x + (p * H + (long long)oh * s) * W + (long long)ow * s;
        float acc = mode == 0 ? -INFINITY : 0.f;
        bool has_nan = false;
        for (int dy = 0; dy < k; ++dy)
            for (int dx = 0; dx < k; ++dx) {
                float v = src[(long long)dy * W + dx];
                if (neg_in) v = -v;
                if (mode == 0) {
                    has_nan |= (v != v);
                    acc = fmaxf(acc, v);
                } else acc += v;
            }
        if (mode == 1) acc /= (float)(k * k);
        else if (has_nan) acc = NAN;  // torch max_pool2d propagates NaN
        y[i] = neg_out ? -acc : acc;
    }
}

struct Strides4 { long long s[4]; };

// area_pixel_compute_source_index (ATen UpSample.h), evaluated in fp32 like PyTorch does
__device__ __forceinline__ void bilinear_src(int dst, int in, int out, bool align, int& i0, int& i1,
                                             float& l1) {
    float src;
    if (align) {
        const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
        src = scale * (float)dst;
    } else {
        const float scale = (float)in / (float)out;
        src = scale * ((float)dst + 0.5f) - 0.5f;
        if (src < 0.f) src = 0.f;
    }
    i0 = (int)src;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

// One thread per 4 consecutive output pixels of a row: the row decomposition (three divisions) and the vertical
// source rows / weight are computed once per 4 outputs, in 32-bit arithmetic when the problem allows (64-bit integer
// division costs ~100 instructions on the GPU and used to dominate this kernel).
template <typename Idx>
__global__ void __launch_bounds__(256)
k_resize_bilinear(const float* __restrict__ x, float* __restrict__ y, int N, int C, int H, int W, int OH,
                  int OW, int align, Strides4 xs, Strides4 ys, float mul, float add, float lo, float hi,
                  int pre, float pre_mul, float pre_lo, float pre_hi) {
    const Idx OWQ = (Idx)((OW + 3) >> 2);
    const Idx total = (Idx)N * (Idx)C * (Idx)OH * OWQ;
    const Idx stride = (Idx)gridDim.x * blockDim.x;
    for (Idx i = (Idx)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const Idx rowi = i / OWQ;
        const int owq = (int)(i - rowi * OWQ);
        const Idx nc = rowi / (Idx)OH;
        const int oh = (int)(rowi - nc * (Idx)OH);
        const Idx n = nc / (Idx)C;
        const int c = (int)(nc - n * (Idx)C);
        int h0, h1;
        float lh;
        bilinear_src(oh, H, OH, align, h0, h1, lh);
        const float* p = x + (long long)n * xs.s[0] + (long long)c * xs.s[1];
        const float* r0 = p + (long long)h0 * xs.s[2];
        const float* r1 = p + (long long)h1 * xs.s[2];
        float* yo = y + (long long)n * ys.s[0] + (long long)c * ys.s[1] + (long long)oh * ys.s[2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ow = owq * 4 + k;
            if (ow >= OW) break;
            int w0, w1;
            float lw;
            bilinear_src(ow, W, OW, align, w0, w1, lw);
            float v00 = r0[w0 * xs.s[3]], v01 = r0[w1 * xs.s[3]];
            float v10 = r1[w0 * xs.s[3]], v11 = r1[w1 * xs.s[3]];
            if (pre) {      // the source is seen through clip(v * pre_mul, pre_lo, pre_hi) (NaN kept), without materialising it
                v00 *= pre_mul; v01 *= pre_mul; v10 *= pre_mul; v11 *= pre_mul;
                if (v00 == v00) v00 = fminf(fmaxf(v00, pre_lo), pre_hi);
                if (v01 == v01) v01 = fminf(fmaxf(v01, pre_lo), pre_hi);
                if (v10 == v10) v10 = fminf(fmaxf(v10, pre_lo), pre_hi);
                if (v11 == v11) v11 = fminf(fmaxf(v11, pre_lo), pre_hi);
            }
            float v = (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);
            v = v * mul + add;
            if (v == v) v = fminf(fmaxf(v, lo), hi);
            yo[ow * ys.s[3]] = v;
        }
    }
}

static void launch_resize_bilinear(const float* d_x, float* d_y, int N, int C, int H, int W, int OH, int OW, int align, Strides4 xs, Strides4 ys,
                                   float mul, float add, float lo, float hi, int pre, float pre_mul, float pre_lo, float pre_hi, cudaStream_t st) {
    const long long items = (long long)N * C * OH * ((OW + 3) / 4);
    const int grid = stream_grid(items, 256, 32);
    if (items < (1ll << 31))
        k_resize_bilinear<unsigned><<<grid, 256, 0, st>>>(d_x, d_y, N, C, H, W, OH, OW, align, xs, ys, mul, add, lo, hi, pre, pre_mul, pre_lo, pre_hi);
    else
        k_resize_bilinear<long long><<<grid, 256, 0, st>>>(d_x, d_y, N, C, H, W, OH, OW, align, xs, ys, mul, add, lo, hi, pre, pre_mul, pre_lo, pre_hi);
}

// one warp per row
__global__ void __launch_bounds__(256)
k_layernorm(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
            float* __restrict__ y, long long rows, int C, float eps) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + row * C;
    float v[32];  // C <= 1024
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int c = lane + 32 * j;
        v[j] = c < C ? xr[c] : 0.f;
        sum += v[j];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const float mean = sum / (float)C;
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int c = lane + 32 * j;
        const float dlt = c < C ? v[j] - mean : 0.f;
        var += dlt * dlt;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
    const float rstd = 1.f / sqrtf(var / (float)C + eps);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int c = lane + 32 * j;
        if (c < C) y[row * C + c] = (v[j] - mean) * rstd * gamma[c] + beta[c];
    }
}

// one thread per (b, n, head); keys/values are a handful of tokens
__global__ void __launch_bounds__(128)
k_attention_small(const float* __restrict__ q, const float* __restrict__ kv, float* __restrict__ out,
                  int B, int N, int C, int heads, int n_kv) {
    const long long total = (long long)B * N * heads;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int h = (int)(i % heads);
    const long long bn = i / heads;
    const long long b = bn / N;
    const int d = C / heads;
    const float* qp = q + bn * C + h * d;
    const float* kvb = kv + b * n_kv * 2 * C;
    const float inv = 1.f / sqrtf((float)C / (float)heads);
    float sc[32];
    float mx = -INFINITY;
    for (int s = 0; s < n_kv; ++s) {
        const float* kp = kvb + (long long)s * 2 * C + h * d;
        float dot = 0.f;
        for (int j = 0; j < d; ++j) dot = fmaf(qp[j], kp[j], dot);
        sc[s] = dot * inv;
        mx = fmaxf(mx, sc[s]);
    }
    float den = 0.f;
    for (int s = 0; s < n_kv; ++s) {
        sc[s] = expf(sc[s] - mx);
        den += sc[s];
    }
    float* op = out + bn * C + h * d;
    for (int j = 0; j < d; ++j) {
        float acc = 0.f;
        for (int s = 0; s < n_kv; ++s) acc = fmaf(sc[s] / den, kvb[(long long)s * 2 * C + C + h * d + j], acc);
        op[j] = acc;
    }
}

struct Dims4 { long long d[4]; };

__global__ void __launch_bounds__(256)
k_map4d(const float* x, Strides4 xs, float* y, Strides4 ys, Dims4 dm, float mul, float div, float add,
        float lo, float hi) {
    const long long total = dm.d[0] * dm.d[1] * dm.d[2] * dm.d[3];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long i3 = i % dm.d[3], r3 = i / dm.d[3];
        const long long i2 = r3 % dm.d[2], r2 = r3 / dm.d[2];
        const long long i1 = r2 % dm.d[1], i0 = r2 / dm.d[1];
        float v = x[i0 * xs.s[0] + i1 * xs.s[1] + i2 * xs.s[2] + i3 * xs.s[3]];
        v = __fadd_rn(__fdiv_rn(__fmul_rn(v, mul), div), add);  // mul = div = 1, add = 0 are exact no-ops
        if (v == v) v = fminf(fmaxf(v, lo), hi);
        y[i0 * ys.s[0] + i1 * ys.s[1] + i2 * ys.s[2] + i3 * ys.s[3]] = v;
    }
}

__global__ void __launch_bounds__(256)
k_pixel_shuffle(const float* __restrict__ x, float* __restrict__ y, int N, int C, int H, int W, int r,
                Strides4 xs, Strides4 ys) {
    const int OH = H * r, OW = W * r;
    const long long total = (long long)N * C * OH * OW;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int ow = (int)(i % OW), oh = (int)((i / OW) % OH);
        const int c = (int)((i / ((long long)OW * OH)) % C);
        const long long n = i / ((long long)OW * OH * C);
        const int ci = c * r * r + (oh % r) * r + (ow % r);
        y[n * ys.s[0] + c * ys.s[1] + oh * ys.s[2] + ow * ys.s[3]] =
            x[n * xs.s[0] + ci * xs.s[1] + (oh / r) * xs.s[2] + (ow / r) * xs.s[3]];
    }
}

__global__ void __launch_bounds__(256)
k_form_input(float* __restrict__ x, float* __restrict__ out, long long n, long long plane, int form_bev,
             float cutoff) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = x[i];
        if (fabsf(v) < cutoff) {  // NaN compares false and stays
            v = 0.f;
            x[i] = 0.f;
        }
        if (form_bev == 0) {
            const float pos = v > 0.f ? v : 0.f;
            const long long fr = i / plane, px = i - fr * plane;
            out[(fr * 2) * plane + px] = pos;
            out[(fr * 2 + 1) * plane + px] = pos;
        } else if (form_bev == 1) {
            out[i] = fabsf(v);
        } else {
            out[i] = (v != 0.f) ? 1.f : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------
// LSTM layer over an unbatched sequence: one persistent CTA, h and c in shared memory,
// W_hh^T [H,4H] streamed from L2 every step (coalesced across the 4H gate rows).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_lstm_seq(const float* __restrict__ gx, const float* __restrict__ whh_t, const float* __restrict__ h0,
           const float* __restrict__ c0, float* __restrict__ hs, float* __restrict__ hT,
           float* __restrict__ cT, int T, int H, int n_seq) {
    extern __shared__ float sm[];
    {   // sequence blockIdx.x of n_seq: gx [T, n_seq, 4H], hs [T, n_seq, H], states [n_seq, H]
        const int sq = blockIdx.x;
        gx += (long long)sq * 4 * H;
        hs += (long long)sq * H;
        if (h0) h0 += (long long)sq * H;
        if (c0) c0 += (long long)sq * H;
        if (hT) hT += (long long)sq * H;
        if (cT) cT += (long long)sq * H;
    }
    const long long gstep = (long long)n_seq * 4 * H, hstep = (long long)n_seq * H;
    float* s_h = sm;            // [H]
    float* s_c = sm + H;        // [H]
    float* s_g = sm + 2 * H;    // [4H]
    const int G = 4 * H;
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        s_h[j] = h0 ? h0[j] : 0.f;
        s_c[j] = c0 ? c0[j] : 0.f;
    }
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        for (int r = threadIdx.x; r < G; r += blockDim.x) {
            float acc = gx[(long long)t * gstep + r];
            const float* wp = whh_t + r;
#pragma unroll 4
            for (int j = 0; j < H; ++j) acc = fmaf(s_h[j], __ldg(wp + (long long)j * G), acc);
            s_g[r] = acc;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < H; j += blockDim.x) {
            const float ig = 1.f / (1.f + expf(-s_g[j]));
            const float fg = 1.f / (1.f + expf(-s_g[H + j]));
            const float gg = tanhf(s_g[2 * H + j]);
            const float og = 1.f / (1.f + expf(-s_g[3 * H + j]));
            const float c = fg * s_c[j] + ig * gg;
            const float h = og * tanhf(c);
            s_c[j] = c;
            s_h[j] = h;
            hs[(long long)t * hstep + j] = h;
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        if (hT) hT[j] = s_h[j];
        if (cT) cT[j] = s_c[j];
    }
}

// LSTM cell update for a few rows (short sequences / batch-1 streaming): gates [n, 4H] (i,f,g,o),
// c [n,H] in -> out, h [n,H] out. The gate GEMV runs on k_linear_smallm across many CTAs.
__global__ void __launch_bounds__(256)
k_lstm_pointwise(const float* __restrict__ gates, const float* __restrict__ c_in, float* __restrict__ c_out,
                 float* __restrict__ h_out, float* __restrict__ h_out2, int n, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * H) return;
    const int r = i / H, j = i - r * H;
    const float* g = gates + (long long)r * 4 * H;
    const float ig = 1.f / (1.f + expf(-g[j])), fg = 1.f / (1.f + expf(-g[H + j]));
    const float gg = tanhf(g[2 * H + j]), og = 1.f / (1.f + expf(-g[3 * H + j]));
    const float c = fg * (c_in ? c_in[i] : 0.f) + ig * gg;
    const float h = og * tanhf(c);
    c_out[i] = c;
    h_out[i] = h;
    if (h_out2) h_out2[i] = h;
}

__global__ void __launch_bounds__(256)
k_convlstm_pointwise(const float* __restrict__ gates, float* __restrict__ c, float* __restrict__ h,
                     int Ch, int P) {
    const int total = Ch * P;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float gi = gates[i], gf = gates[total + i], go = gates[2 * total + i], gg = gates[3 * total + i];
    const float iv = 1.f / (1.f + expf(-gi)), fv = 1.f / (1.f + expf(-gf)), ov = 1.f / (1.f + expf(-go));
    const float cn = fv * c[i] + iv * tanhf(gg);
    c[i] = cn;
    h[i] = ov * tanhf(cn);
}

__global__ void k_velpred_unit(const float* __restrict__ y, float* __restrict__ out, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float v = y[i];
    float rad = 1.f - v * v;
    if (rad == rad) rad = fminf(fmaxf(rad, 0.f), 1.f);
    out[i * 3 + 0] = sqrtf(rad);
    out[i * 3 + 1] = v;
    out[i * 3 + 2] = 0.f;
}

}  // namespace evfly

using namespace evfly;

static inline int ew_grid(long long n) { return stream_grid(n, 256 * 4, 16); }

extern "C" int evfly_conv2d_f32(const evfly_conv2d_args* p, void* stream) {
    EVFLY_REQUIRE(p, "conv2d_f32: null args");
    const evfly_conv2d_args a = *p;
    EVFLY_REQUIRE(a.x && a.w && a.y, "conv2d_f32: null tensor");
    EVFLY_REQUIRE(a.N >= 0 && a.Cin > 0 && a.H > 0 && a.W > 0 && a.Cout > 0 && a.KH > 0 && a.KW > 0 && a.stride > 0 && a.pad >= 0 && a.groups > 0,
                  "conv2d_f32: bad shape");
    EVFLY_REQUIRE(a.Cin % a.groups == 0 && a.Cout % a.groups == 0, "conv2d_f32: channels not divisible by groups");
    EVFLY_REQUIRE((a.post_scale == nullptr) == (a.post_shift == nullptr), "conv2d_f32: post_scale/post_shift go together");
    EVFLY_REQUIRE(a.act >= 0 && a.act <= EVFLY_ACT_SIGMOID, "conv2d_f32: bad activation %d", a.act);
    const int OH = (a.H + 2 * a.pad - a.KH) / a.stride + 1, OW = (a.W + 2 * a.pad - a.KW) / a.stride + 1;
    EVFLY_REQUIRE(OH > 0 && OW > 0, "conv2d_f32: empty output");
    if (a.N == 0) return EVFLY_OK;
    const long long M = (long long)a.N * OH * OW;
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups;
    const int K = cin_g * a.KH * a.KW;
    if (M * a.Cout <= 32768 && K <= 2048 && ceil_div(M, CBM) * ceil_div(cout_g, CBN) * a.groups < 64) {
        k_conv2d_small<<<(unsigned)ceil_div(M * a.Cout, 8), 256, 0, (cudaStream_t)stream>>>(a, OH, OW, M, K, cin_g, cout_g);
        EVFLY_LAUNCHED();
        return EVFLY_OK;
    }
    const long long gx = ceil_div(M, CBM);
    EVFLY_REQUIRE(gx < (1ll << 31) && a.groups < 65536, "conv2d_f32: problem too large for one grid");
    dim3 grid((unsigned)gx, (unsigned)ceil_div(cout_g, CBN), (unsigned)a.groups);
    k_conv2d_f32<<<grid, 256, 0, (cudaStream_t)stream>>>(a, OH, OW, M, K, cin_g, cout_g);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_linear_smallm_f32(const float* d_x, int64_t x_ld, const float* d_w, const float* d_bias, const float* d_res,
                                       int64_t res_ld, float* d_y, int64_t y_ld, int M, int N, int K, int act, void* stream) {
    EVFLY_REQUIRE(d_x && d_w && d_y && M > 0 && M <= kSmallM && N > 0 && K > 0 && (size_t)M * K * 4 <= 160 * 1024, "linear_smallm_f32: bad argument (M <= 8)");
    EVFLY_REQUIRE(act >= 0 && act <= EVFLY_ACT_SIGMOID, "linear_smallm_f32: bad activation");
    const size_t smem = (size_t)M * K * sizeof(float);
    EVFLY_SMEM_ATTR(160 * 1024, k_linear_smallm);
    k_linear_smallm<<<(unsigned)ceil_div(N, 8), 256, smem, (cudaStream_t)stream>>>(d_x, x_ld, d_w, d_bias, d_res, res_ld, d_y, y_ld, M, N, K, act);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_pool2d_f32(const float* d_x, float* d_y, int64_t planes, int H, int W, int k,
                                int stride, int mode, int negate_in, int negate_out, void* stream) {
    EVFLY_REQUIRE(d_x && d_y && planes >= 0 && H >= k && W >= k && k > 0 && stride > 0 && (mode == 0 || mode == 1), "pool2d_f32: bad argument");
    if (planes == 0) return EVFLY_OK;
    const int OH = (H - k) / stride + 1, OW = (W - k) / stride + 1;
    k_pool2d<<<ew_grid(planes * OH * OW), 256, 0, (cudaStream_t)stream>>>(d_x, d_y, planes, H, W, OH, OW, k, stride, mode, negate_in, negate_out);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

static inline Strides4 to_strides(const int64_t* p) {
    Strides4 s;
    for (int i = 0; i < 4; ++i) s.s[i] = p[i];
    return s;
}

extern "C" int evfly_resize_bilinear_f32(const float* d_x, const int64_t* xs, float* d_y, const int64_t* ys,
                                         int N, int C, int H, int W, int OH, int OW, int align_corners,
                                         float mul, float add, float lo, float hi, void* stream) {
    EVFLY_REQUIRE(d_x && d_y && xs && ys && N >= 0 && C > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, "resize_bilinear_f32: bad argument");
    if (N == 0) return EVFLY_OK;
    launch_resize_bilinear(d_x, d_y, N, C, H, W, OH, OW, align_corners, to_strides(xs), to_strides(ys), mul, add, lo, hi, 0, 1.f, 0.f, 0.f,
                           (cudaStream_t)stream);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_resize_bilinear_premap_f32(const float* d_x, const int64_t* xs, float* d_y, const int64_t* ys, int N, int C, int H, int W,
                                               int OH, int OW, int align_corners, float pre_mul, float pre_lo, float pre_hi, void* stream) {
    EVFLY_REQUIRE(d_x && d_y && xs && ys && N >= 0 && C > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, "resize_bilinear_premap_f32: bad argument");
    if (N == 0) return EVFLY_OK;
    launch_resize_bilinear(d_x, d_y, N, C, H, W, OH, OW, align_corners, to_strides(xs), to_strides(ys), 1.f, 0.f, -INFINITY, INFINITY, 1, pre_mul,
                           pre_lo, pre_hi, (cudaStream_t)stream);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_layernorm_f32(const float* d_x, const float* d_gamma, const float* d_beta, float* d_y,
                                   int64_t rows, int C, float eps, void* stream) {
    EVFLY_REQUIRE(d_x && d_gamma && d_beta && d_y && rows >= 0 && C > 0 && C <= 1024, "layernorm_f32: bad argument (C <= 1024)");
    if (rows == 0) return EVFLY_OK;
    k_layernorm<<<(unsigned)ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(d_x, d_gamma, d_beta, d_y, rows, C, eps);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_attention_small_f32(const float* d_q, const float* d_kv, float* d_out, int B, int N,
                                         int C, int heads, int n_kv, void* stream) {
    EVFLY_REQUIRE(d_q && d_kv && d_out && B >= 0 && N > 0 && C > 0 && heads > 0 && C % heads == 0 && n_kv > 0 && n_kv <= 32,
                  "attention_small_f32: bad argument (n_kv <= 32)");
    if (B == 0) return EVFLY_OK;
    const long long total = (long long)B * N * heads;
    k_attention_small<<<(unsigned)ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>(d_q, d_kv, d_out, B, N, C, heads, n_kv);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_map4d_f32(const float* d_x, const int64_t* xs, float* d_y, const int64_t* ys,
                               const int64_t* dims, float mul, float div, float add, float lo, float hi,
                               void* stream) {
    EVFLY_REQUIRE(d_x && d_y && xs && ys && dims, "map4d_f32: null pointer");
    Dims4 dm;
    long long total = 1;
    for (int i = 0; i < 4; ++i) {
        EVFLY_REQUIRE(dims[i] >= 0, "map4d_f32: negative dimension");
        dm.d[i] = dims[i];
        total *= dims[i];
    }
    EVFLY_REQUIRE(div != 0.f, "map4d_f32: div must be non-zero");
    if (total == 0) return EVFLY_OK;
    k_map4d<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(d_x, to_strides(xs), d_y, to_strides(ys), dm, mul, div, add, lo, hi);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_pixel_shuffle_f32(const float* d_x, const int64_t* xs, float* d_y, const int64_t* ys,
                                       int N, int C, int H, int W, int r, void* stream) {
    EVFLY_REQUIRE(d_x && d_y && xs && ys && N >= 0 && C > 0 && H > 0 && W > 0 && r > 0, "pixel_shuffle_f32: bad argument");
    if (N == 0) return EVFLY_OK;
    k_pixel_shuffle<<<ew_grid((long long)N * C * H * W * r * r), 256, 0, (cudaStream_t)stream>>>(
        d_x, d_y, N, C, H, W, r, to_strides(xs), to_strides(ys));
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_form_input_f32(float* d_x, float* d_out, int64_t n, int64_t plane, int form_bev,
                                    float cutoff, void* stream) {
    EVFLY_REQUIRE(d_x && d_out && n >= 0 && plane > 0 && form_bev >= 0 && form_bev <= 2, "form_input_f32: bad argument");
    if (n == 0) return EVFLY_OK;
    k_form_input<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_x, d_out, n, plane, form_bev, cutoff);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_lstm_seq_f32(const float* d_gx, const float* d_whh_t, const float* d_h0,
                                  const float* d_c0, float* d_hs, float* d_hT, float* d_cT, int T, int H,
                                  int n_seq, void* stream) {
    EVFLY_REQUIRE(d_gx && d_whh_t && d_hs && T >= 0 && H > 0 && H <= 2048 && n_seq > 0, "lstm_seq_f32: bad argument");
    const size_t smem = (size_t)6 * H * sizeof(float);
    const int threads = 4 * H >= 1024 ? 1024 : ((4 * H + 31) / 32) * 32;
    k_lstm_seq<<<n_seq, threads, smem, (cudaStream_t)stream>>>(d_gx, d_whh_t, d_h0, d_c0, d_hs, d_hT, d_cT, T, H, n_seq);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_lstm_pointwise_f32(const float* d_gates, const float* d_c_in, float* d_c_out, float* d_h_out, float* d_h_out2,
                                        int n, int H, void* stream) {
    EVFLY_REQUIRE(d_gates && d_c_out && d_h_out && n > 0 && H > 0, "lstm_pointwise_f32: bad argument");
    k_lstm_pointwise<<<(n * H + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_gates, d_c_in, d_c_out, d_h_out, d_h_out2, n, H);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_convlstm_pointwise_f32(const float* d_gates, float* d_c, float* d_h_out, int Ch, int P,
                                            void* stream) {
    EVFLY_REQUIRE(d_gates && d_c && d_h_out && Ch > 0 && P > 0, "convlstm_pointwise_f32: bad argument");
    k_convlstm_pointwise<<<(Ch * P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_gates, d_c, d_h_out, Ch, P);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_velpred_unit_f32(const float* d_y, float* d_out, int N, void* stream) {
    EVFLY_REQUIRE(d_y && d_out && N >= 0, "velpred_unit_f32: bad argument");
    if (N == 0) return EVFLY_OK;
    k_velpred_unit<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_y, d_out, N);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}
