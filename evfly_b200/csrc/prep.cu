// prep.cu -- the formats either side of the accumulation kernels:
//   * packers: float64 [t,x,y,p] rows (form_eventframe's argument) / SoA int64 streams
//     (to_events.py) -> 16-byte evfly_event records, with the reference's masks applied;
//   * L2 frame normalisation: u8 decode + centre crop (evfly_ros/run.py:334-350) and the
//     97th-percentile scale / clip / min-cutoff (run.py:250-253, dataloading.py:518-533),
//     the quantile found exactly by radix select instead of a full sort.
#include "common.cuh"

namespace evfly {

// ---------------------------------------------------------------------------------------
// packers
// ---------------------------------------------------------------------------------------
// np.histogram2d(range=[[0,W],[0,H]], bins=(W,H)) -> unit bins: floor(c) for 0 <= c < n,
// c == n lands in the last bin, everything else (and NaN) is an outlier.
__device__ __forceinline__ int hist_bin(double c, int nbins) {
    if (!(c >= 0.0) || c > (double)nbins) return -1;
    int b = (int)c;  // trunc == floor for c >= 0
    if (b == nbins) b = nbins - 1;
    return b;
}

__device__ __forceinline__ unsigned pol_class_f64(double p, int pol_mode) {
    if (p > 0.0) return EVFLY_POL_POS;
    if (pol_mode == EVFLY_NEG_IS_ZERO) return p == 0.0 ? EVFLY_POL_NEG : EVFLY_POL_SKIP;
    return p < 0.0 ? EVFLY_POL_NEG : EVFLY_POL_SKIP;
}

__device__ __forceinline__ uint4 make_record(int bx, int by, double t, unsigned pol) {
    // t is in ns; records carry floor(t) split into sec / nsec (negative or huge t is only
    // meaningful to the time masks, which are applied on the float value before this point)
    long long tn = 0;
    if (t >= 0.0 && t < 9.2e18) tn = (long long)floor(t);
    unsigned long long sec = (unsigned long long)tn / 1000000000ull;
    if (sec > 0xffffffffull) sec = 0xffffffffull;
    const unsigned nsec = (unsigned)((unsigned long long)tn - sec * 1000000000ull);
    uint4 r;
    r.x = ((unsigned)by << 16) | (unsigned)bx;
    r.y = (unsigned)sec;
    r.z = nsec;
    r.w = pol;
    return r;
}

// rows that pass the `t >= t_lo` mask get an ordered rank (N-mode of form_eventframe)
__global__ void __launch_bounds__(1024)
k_pack_count(const double* __restrict__ rows, int64_t n, double t_lo, unsigned long long* __restrict__ block_counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool pass = (i < n) && (rows[i * 4] >= t_lo);
    const int c = __syncthreads_count(pass);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (unsigned long long)c;
}

// single CTA: exclusive scan of the per-block counts (in place); total -> block_counts[nblocks]
__global__ void __launch_bounds__(1024)
k_pack_scan(unsigned long long* __restrict__ block_counts, int64_t nblocks) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nblocks; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const unsigned long long v = (i < nblocks) ? block_counts[i] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane_id() >= d) x += y;
        }
        if (lane_id() == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long w = s_warp[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, w, d);
                if (lane_id() >= d) w += y;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const unsigned long long warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0ull;
        const unsigned long long incl = s_carry + warp_off + x;
        if (i < nblocks) block_counts[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_counts[nblocks] = s_carry;
}

__global__ void __launch_bounds__(1024)
k_pack_f64(const double* __restrict__ rows, int64_t n, int H, int W, int pol_mode, int use_time,
           double t_lo, double t_hi, int64_t max_events,
           const unsigned long long* __restrict__ block_offsets, uint4* __restrict__ out,
           double* __restrict__ last_kept_t) {
    __shared__ int s_warp[32];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double t = 0.0, x = -1.0, y = -1.0, p = 0.0;
    if (i < n) {
        // 32-byte row: two 16-byte loads
        const double2 a = reinterpret_cast<const double2*>(rows)[i * 2];
        const double2 b = reinterpret_cast<const double2*>(rows)[i * 2 + 1];
        t = a.x; x = a.y; y = b.x; p = b.y;
    }
    bool keep = i < n;
    if (max_events >= 0) {
        // ordered rank among rows with t >= t_lo (ev_utils.py:131: view_events[t >= t0][:N])
        const bool pass = keep && (t >= t_lo);
        const unsigned ball = __ballot_sync(0xffffffffu, pass);
        const int in_warp = __popc(ball & ((1u << lane_id()) - 1u));
        if (lane_id() == 0) s_warp[threadIdx.x >> 5] = __popc(ball);
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
            const int orig = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int yv = __shfl_up_sync(0xffffffffu, w, d);
                if (lane_id() >= d) w += yv;
            }
            s_warp[threadIdx.x] = w - orig;  // exclusive
        }
        __syncthreads();
        const unsigned long long rank = block_offsets[blockIdx.x] + (unsigned long long)(s_warp[threadIdx.x >> 5] + in_warp);
        const unsigned long long total = block_offsets[gridDim.x];
        keep = pass && rank < (unsigned long long)max_events;
        const unsigned long long last = (total < (unsigned long long)max_events ? total : (unsigned long long)max_events);
        if (pass && last_kept_t && last > 0 && rank == last - 1) *last_kept_t = t;
    } else if (use_time) {
        keep = keep && (t >= t_lo) && (t < t_hi);
    }
    if (i >= n) return;
    const int bx = hist_bin(x, W), by = hist_bin(y, H);
    unsigned pol = pol_class_f64(p, pol_mode);
    if (!keep || bx < 0 || by < 0) pol = EVFLY_POL_SKIP;
    out[i] = make_record(bx < 0 ? 0 : bx, by < 0 ? 0 : by, t, pol);
}

__global__ void __launch_bounds__(256)
k_pack_soa(const int64_t* __restrict__ xs, const int64_t* __restrict__ ys,
           const int64_t* __restrict__ ts, const int64_t* __restrict__ ps, int64_t n, int H, int W,
           int pol_mode, uint4* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t x = xs[i], y = ys[i], t = ts[i], p = ps[i];
        // integer coordinates through np.histogram2d: 0..W-1 -> own bin, W -> bin W-1
        int bx = (x >= 0 && x <= W) ? (int)(x == W ? W - 1 : x) : -1;
        int by = (y >= 0 && y <= H) ? (int)(y == H ? H - 1 : y) : -1;
        unsigned pol;
        if (p > 0) pol = EVFLY_POL_POS;
        else if (pol_mode == EVFLY_NEG_IS_ZERO) pol = (p == 0) ? EVFLY_POL_NEG : EVFLY_POL_SKIP;
        else pol = (p < 0) ? EVFLY_POL_NEG : EVFLY_POL_SKIP;
        if (bx < 0 || by < 0) { pol = EVFLY_POL_SKIP; bx = 0; by = 0; }
        // exact integer time (a double would round above 2^53)
        const unsigned long long tn = t < 0 ? 0ull : (unsigned long long)t;
        unsigned long long sec = tn / 1000000000ull;
        if (sec > 0xffffffffull) sec = 0xffffffffull;
        uint4 r;
        r.x = ((unsigned)by << 16) | (unsigned)bx;
        r.y = (unsigned)sec;
        r.z = (unsigned)(tn - sec * 1000000000ull);
        r.w = pol;
        out[i] = r;
    }
}

// ---------------------------------------------------------------------------------------
// decode + centre crop
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_decode_crop(const uint8_t* __restrict__ u8, const int* __restrict__ counts, int N, int H, int W,
              int h, int w, int r0, int c0, float scale, float* __restrict__ out) {
    const int64_t total = (int64_t)N * h * w;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int col = (int)(i % w);
        const int row = (int)((i / w) % h);
        const int nfr = (int)(i / ((int64_t)w * h));
        const int64_t src = (int64_t)(row + r0) * W + (col + c0);
        float d;
        if (u8) {
            d = (float)u8[(int64_t)nfr * H * W + src] - 128.0f;  // run.py:334-335
        } else {
            const int* c = counts + (int64_t)nfr * 2 * H * W;
            d = (float)(c[(int64_t)H * W + src] - c[src]);
        }
        out[i] = __fmul_rn(d, scale);  // run.py:336
    }
}

// ---------------------------------------------------------------------------------------
// exact per-frame quantile of |x| by 3-pass radix select, then scale / clip / cutoff
// ---------------------------------------------------------------------------------------
// One CTA per frame. |x| of a finite fp32 orders like its 31-bit pattern, so the k-th smallest
// is found digit by digit (11 + 10 + 10 bits) with shared-memory histograms; the frame
// (360 KB at 260x346) is re-read from L2. NaN patterns sort above +inf, like torch's sort.
constexpr int kSelThreads = 1024;
constexpr int kSelBins = 2048;

struct SelectResult {
    unsigned bits;               // pattern of the k-th smallest |x|
    unsigned long long n_le;     // number of elements <= that value
};

__device__ SelectResult radix_select(const float* x, int64_t n, unsigned long long k,
                                     unsigned* s_hist, unsigned long long* s_misc) {
    unsigned prefix = 0, mask = 0;
    unsigned long long below = 0;  // elements strictly below the current prefix bucket
    const int shifts[3] = {20, 10, 0};
    const int widths[3] = {11, 10, 10};
    unsigned long long n_le = 0;
    for (int pass = 0; pass < 3; ++pass) {
        const int sh = shifts[pass];
        const unsigned nb = 1u << widths[pass];
        for (unsigned b = threadIdx.x; b < nb; b += blockDim.x) s_hist[b] = 0;
        __syncthreads();
        // Event frames are ~65 % zeros: every zero lands in bin 0 of every pass, which would serialise the
        // shared-memory atomics on one bank, so bin 0 is counted in a register and added once per thread.
        unsigned zero_bin = 0;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned u = __float_as_uint(x[i]) & 0x7fffffffu;
            if ((u & mask) == prefix) {
                const unsigned key = (u >> sh) & (nb - 1);
                if (key == 0) ++zero_bin; else atomicAdd(&s_hist[key], 1u);
            }
        }
        if (zero_bin) atomicAdd(&s_hist[0], zero_bin);
        __syncthreads();
        // parallel search of the bin that holds rank k: every thread owns nb/1024 consecutive bins, block-wide
        // exclusive scan of the per-thread sums (warp shuffles + one warp over the 32 warp totals)
        {
            const unsigned per = nb / kSelThreads;            // 2 (11-bit pass) or 1
            const unsigned b0 = threadIdx.x * per;
            const unsigned c0 = s_hist[b0], c1 = per == 2 ? s_hist[b0 + 1] : 0u;
            unsigned incl = c0 + c1;
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += y;
            }
            __shared__ unsigned s_wsum[32];
            if (lane == 31) s_wsum[wid] = incl;
            __syncthreads();
            if (wid == 0) {
                unsigned wv = s_wsum[lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned y = __shfl_up_sync(0xffffffffu, wv, d);
                    if (lane >= d) wv += y;
                }
                s_wsum[lane] = wv;       // inclusive warp totals
            }
            __syncthreads();
            const unsigned long long excl = below + (wid ? s_wsum[wid - 1] : 0u) + (incl - c0 - c1);
            // k falls into one of this thread's bins?
            if (k >= excl && k < excl + c0) { s_misc[0] = b0; s_misc[1] = excl; s_misc[2] = excl + c0; }
            else if (per == 2 && k >= excl + c0 && k < excl + c0 + c1) { s_misc[0] = b0 + 1; s_misc[1] = excl + c0; s_misc[2] = excl + c0 + c1; }
        }
        __syncthreads();
        const unsigned b = (unsigned)s_misc[0];
        below = s_misc[1];
        n_le = s_misc[2];
        prefix |= b << sh;
        mask |= (nb - 1) << sh;
        __syncthreads();
    }
    SelectResult r;
    r.bits = prefix;
    r.n_le = n_le;
    return r;
}

__global__ void __launch_bounds__(kSelThreads)
k_quantile_scale_clip(const float* x, int64_t elems, float qfrac, float lo, float hi,
                      float cutoff, float* out, float* __restrict__ qout) {  // x may alias out
    __shared__ unsigned s_hist[kSelBins];
    __shared__ unsigned long long s_misc[4];
    __shared__ float s_q;
    const float* xf = x + (int64_t)blockIdx.x * elems;
    float* of = out + (int64_t)blockIdx.x * elems;

    // torch.quantile (aten/native/Sorting.cpp quantile_compute): rank = q * (n-1) evaluated in
    // the INPUT dtype (fp32), below = floor, above = ceil, result = lerp(v_below, v_above, frac)
    const float rank = __fmul_rn(qfrac, (float)(elems - 1));
    const float rb = floorf(rank), ra = ceilf(rank);
    const unsigned long long kb = (unsigned long long)rb, ka = (unsigned long long)ra;
    const float wgt = __fsub_rn(rank, rb);

    const SelectResult sb = radix_select(xf, elems, kb, s_hist, s_misc);
    unsigned bits_a = sb.bits;
    if (ka != kb && ka >= sb.n_le) {  // the next order statistic is a different value
        const SelectResult sa = radix_select(xf, elems, ka, s_hist, s_misc);
        bits_a = sa.bits;
    }
    if (threadIdx.x == 0) {
        const float vb = __uint_as_float(sb.bits), va = __uint_as_float(bits_a);
        // at::lerp: w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
        const float diff = __fsub_rn(va, vb);
        const float q = (wgt < 0.5f) ? __fadd_rn(vb, __fmul_rn(wgt, diff))
                                     : __fsub_rn(va, __fmul_rn(diff, __fsub_rn(1.0f, wgt)));
        s_q = q;
        if (qout) qout[blockIdx.x] = q;
    }
    __syncthreads();
    const float q = s_q;
    for (int64_t i = threadIdx.x; i < elems; i += blockDim.x) {
        float v = __fdiv_rn(xf[i], q);       // x / q: 0/0 -> NaN, c/0 -> +-inf (reference F8b)
        if (v == v) v = fminf(fmaxf(v, lo), hi);  // torch.clip propagates NaN; fminf would not
        if (fabsf(v) < cutoff) v = 0.0f;      // NaN compares false, stays NaN like the reference
        of[i] = v;
    }
}

// ---------------------------------------------------------------------------------------
// decode_crop(counts) + quantile_scale_clip in one kernel for frames that come from integer counts:
// |frame| = 0.2f * |n+ - n-| is monotonic in the integer |n+ - n-|, so the two order statistics torch.quantile
// interpolates between are found EXACTLY from one integer histogram pass (2047 levels per pass; hotter pixels take
// another pass with the base moved up), instead of three radix passes over a materialised fp32 frame. The arithmetic
// of the result (fp32 rank, lerp, x / q, clip, cutoff) is the same instruction sequence as the two-kernel path, so
// the output is bit-identical to it (tests/test_accumulate_gpu.py).
// ---------------------------------------------------------------------------------------
constexpr unsigned kLevels = kSelBins - 1;      // bins 0..2046 are exact levels, bin 2047 collects everything above

__global__ void __launch_bounds__(kSelThreads)
k_counts_normalise(const int* __restrict__ counts, int H, int W, int h, int w, int r0, int c0, float scale, float qfrac, float lo,
                   float hi, float cutoff, float* __restrict__ out, float* __restrict__ qout) {
    __shared__ unsigned s_hist[kSelBins];
    __shared__ unsigned s_wsum[32];
    __shared__ unsigned s_level[2];          // |n+ - n-| of the order statistics kb, ka
    __shared__ int s_found[2];
    __shared__ float s_q;
    const int* neg = counts + (int64_t)blockIdx.x * 2 * H * W;
    const int* pos = neg + (int64_t)H * W;
    const int64_t elems = (int64_t)h * w;
    float* of = out + (int64_t)blockIdx.x * elems;
    const float rank = __fmul_rn(qfrac, (float)(elems - 1));
    const float rb = floorf(rank), ra = ceilf(rank);
    const unsigned long long kq[2] = {(unsigned long long)rb, (unsigned long long)ra};
    const float wgt = __fsub_rn(rank, rb);
    if (threadIdx.x < 2) s_found[threadIdx.x] = 0;
    __syncthreads();
    unsigned base = 0;
    unsigned long long below = 0;            // elements with |d| < base
    for (;;) {
        // which ranks earlier passes already resolved: read here, a barrier away from this pass's writes below
        const bool fnd[2] = {s_found[0] != 0, s_found[1] != 0};
        for (unsigned b = threadIdx.x; b < kSelBins; b += blockDim.x) s_hist[b] = 0;
        __syncthreads();
        unsigned zero_bin = 0;               // most pixels of an event frame are empty: count level `base` in a register
        for (int i = threadIdx.x; i < (int)elems; i += blockDim.x) {      // 32-bit index math (elems < 2^24): 64-bit division is ~100 instructions
            const int row = i / w, col = i - row * w;
            const int src = (row + r0) * W + (col + c0);
            const int d = pos[src] - neg[src];
            const unsigned a = (unsigned)(d < 0 ? -d : d);
            if (a >= base) {
                const unsigned key = min(a - base, kLevels);
                if (key == 0) ++zero_bin; else atomicAdd(&s_hist[key], 1u);
            }
        }
        if (zero_bin) atomicAdd(&s_hist[0], zero_bin);
        __syncthreads();
        // block-wide exclusive scan, two bins per thread
        const unsigned b0 = threadIdx.x * 2;
        const unsigned c0v = s_hist[b0], c1v = s_hist[b0 + 1];
        unsigned incl = c0v + c1v;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) s_wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned wv = s_wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, wv, d);
                if (lane >= d) wv += y;
            }
            s_wsum[lane] = wv;
        }
        __syncthreads();
        const unsigned long long excl = below + (wid ? s_wsum[wid - 1] : 0u) + (incl - c0v - c1v);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (!fnd[j]) {
                if (kq[j] >= excl && kq[j] < excl + c0v && b0 < kLevels) { s_level[j] = base + b0; s_found[j] = 1; }
                else if (kq[j] >= excl + c0v && kq[j] < excl + c0v + c1v && b0 + 1 < kLevels) { s_level[j] = base + b0 + 1; s_found[j] = 1; }
            }
        }
        const unsigned in_levels = s_wsum[31] - s_hist[kLevels];     // elements with base <= |d| < base + kLevels
        __syncthreads();
        if ((s_found[0] && s_found[1]) || base > (1u << 30)) break;      // (the bound only guards against garbage counts)
        below += in_levels;
        base += kLevels;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float vb = __fmul_rn((float)s_level[0], scale), va = __fmul_rn((float)s_level[1], scale);
        const float diff = __fsub_rn(va, vb);      // at::lerp, as in k_quantile_scale_clip
        const float q = (wgt < 0.5f) ? __fadd_rn(vb, __fmul_rn(wgt, diff)) : __fsub_rn(va, __fmul_rn(diff, __fsub_rn(1.0f, wgt)));
        s_q = q;
        if (qout) qout[blockIdx.x] = q;
    }
    __syncthreads();
    const float q = s_q;
    for (int i = threadIdx.x; i < (int)elems; i += blockDim.x) {
        const int row = i / w, col = i - row * w;
        const int src = (row + r0) * W + (col + c0);
        float v = __fdiv_rn(__fmul_rn((float)(pos[src] - neg[src]), scale), q);
        if (v == v) v = fminf(fmaxf(v, lo), hi);
        if (fabsf(v) < cutoff) v = 0.0f;
        of[i] = v;
    }
}

}  // namespace evfly

using namespace evfly;

extern "C" int evfly_pack_events_f64(const double* d_rows, int64_t n, int H, int W, int pol_mode,
                                     int use_time, double t_lo, double t_hi, int64_t max_events,
                                     evfly_event* d_out, double* d_last_kept_t, void* d_scan_ws,
                                     void* stream) {
    EVFLY_REQUIRE(n >= 0 && H > 0 && W > 0 && H <= 65535 && W <= 65535, "pack_events_f64: bad n/H/W");
    EVFLY_REQUIRE(pol_mode == EVFLY_NEG_IS_ZERO || pol_mode == EVFLY_NEG_IS_NEGATIVE, "pack_events_f64: bad pol_mode");
    if (n == 0) return EVFLY_OK;
    EVFLY_REQUIRE(d_rows && d_out, "pack_events_f64: null pointer");
    EVFLY_REQUIRE(max_events < 0 || d_scan_ws, "pack_events_f64: N-mode needs d_scan_ws");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nblocks = ceil_div(n, 1024);
    EVFLY_REQUIRE(nblocks < (1ll << 31), "pack_events_f64: n too large");
    unsigned long long* ws = reinterpret_cast<unsigned long long*>(d_scan_ws);
    if (max_events >= 0) {
        k_pack_count<<<(unsigned)nblocks, 1024, 0, st>>>(d_rows, n, t_lo, ws);
        EVFLY_LAUNCHED();
        k_pack_scan<<<1, 1024, 0, st>>>(ws, nblocks);
        EVFLY_LAUNCHED();
    }
    k_pack_f64<<<(unsigned)nblocks, 1024, 0, st>>>(d_rows, n, H, W, pol_mode, use_time, t_lo, t_hi,
                                                   max_events, ws, reinterpret_cast<uint4*>(d_out),
                                                   d_last_kept_t);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_pack_events_soa(const int64_t* d_x, const int64_t* d_y, const int64_t* d_t_ns,
                                     const int64_t* d_p, int64_t n, int H, int W, int pol_mode,
                                     evfly_event* d_out, void* stream) {
    EVFLY_REQUIRE(n >= 0 && H > 0 && W > 0 && H <= 65535 && W <= 65535, "pack_events_soa: bad n/H/W");
    EVFLY_REQUIRE(pol_mode == EVFLY_NEG_IS_ZERO || pol_mode == EVFLY_NEG_IS_NEGATIVE, "pack_events_soa: bad pol_mode");
    if (n == 0) return EVFLY_OK;
    EVFLY_REQUIRE(d_x && d_y && d_t_ns && d_p && d_out, "pack_events_soa: null pointer");
    k_pack_soa<<<stream_grid(n, 256 * 4, 8), 256, 0, (cudaStream_t)stream>>>(
        d_x, d_y, d_t_ns, d_p, n, H, W, pol_mode, reinterpret_cast<uint4*>(d_out));
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_decode_crop(const uint8_t* d_u8, const int32_t* d_counts, int N, int H, int W,
                                 int h, int w, float scale, float* d_out, void* stream) {
    EVFLY_REQUIRE(N >= 0 && H > 0 && W > 0 && h > 0 && w > 0 && h <= H && w <= W, "decode_crop: bad shape");
    EVFLY_REQUIRE((d_u8 != nullptr) != (d_counts != nullptr), "decode_crop: give exactly one of d_u8 / d_counts");
    EVFLY_REQUIRE(d_out, "decode_crop: null output");
    if (N == 0) return EVFLY_OK;
    // run.py:347-348: rows H//2 - h//2 : H//2 + h//2 (so h, w are expected even)
    EVFLY_REQUIRE((h == H || h % 2 == 0) && (w == W || w % 2 == 0), "decode_crop: crop sizes must be even");
    const int r0 = (h == H) ? 0 : H / 2 - h / 2, c0 = (w == W) ? 0 : W / 2 - w / 2;
    const int64_t total = (int64_t)N * h * w;
    k_decode_crop<<<stream_grid(total, 256 * 4, 8), 256, 0, (cudaStream_t)stream>>>(
        d_u8, d_counts, N, H, W, h, w, r0, c0, scale, d_out);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_quantile_scale_clip(const float* d_x, int N, int64_t elems_per_frame,
                                         float qfrac, float lo, float hi, float cutoff,
                                         float* d_out, float* d_q, void* stream) {
    EVFLY_REQUIRE(N >= 0 && elems_per_frame > 0 && elems_per_frame < (1ll << 24),
                  "quantile_scale_clip: bad N / elems (fp32 rank arithmetic needs elems < 2^24)");
    EVFLY_REQUIRE(qfrac >= 0.f && qfrac <= 1.f, "quantile_scale_clip: q must be in [0,1]");
    EVFLY_REQUIRE(d_x && d_out, "quantile_scale_clip: null pointer");
    if (N == 0) return EVFLY_OK;
    k_quantile_scale_clip<<<N, kSelThreads, 0, (cudaStream_t)stream>>>(
        d_x, elems_per_frame, qfrac, lo, hi, cutoff, d_out, d_q);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_counts_normalise(const int32_t* d_counts, int N, int H, int W, int h, int w, float scale, float qfrac, float lo,
                                      float hi, float cutoff, float* d_out, float* d_q, void* stream) {
    EVFLY_REQUIRE(N >= 0 && H > 0 && W > 0 && h > 0 && w > 0 && h <= H && w <= W, "counts_normalise: bad shape");
    EVFLY_REQUIRE((h == H || h % 2 == 0) && (w == W || w % 2 == 0), "counts_normalise: crop sizes must be even");
    EVFLY_REQUIRE((int64_t)h * w < (1ll << 24) && (int64_t)H * W < (1ll << 31), "counts_normalise: fp32 rank arithmetic needs h*w < 2^24 (and H*W < 2^31)");
    EVFLY_REQUIRE(qfrac >= 0.f && qfrac <= 1.f && scale > 0.f, "counts_normalise: q must be in [0,1], scale > 0");
    EVFLY_REQUIRE(d_counts && d_out, "counts_normalise: null pointer");
    if (N == 0) return EVFLY_OK;
    const int r0 = (h == H) ? 0 : H / 2 - h / 2, c0 = (w == W) ? 0 : W / 2 - w / 2;
    k_counts_normalise<<<N, kSelThreads, 0, (cudaStream_t)stream>>>(d_counts, H, W, h, w, r0, c0, scale, qfrac, lo, hi, cutoff, d_out, d_q);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}
