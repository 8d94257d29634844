// next_rows.cu -- the "next" rows either side of the hot path:
//  N2  difflog event approximation used by the simulator front end
//      (envtest/ros/run_competition.py:603-635, utils/to_events.py:417-439): quantise the log-intensity
//      difference by the contrast thresholds, in float64 like the reference's numpy code, with numpy's
//      floor_divide semantics (npy_divmod) reproduced exactly;
//  N1  the min-cutoff pass of the dataset normalisation (learner/dataloading.py:531-533).
#include "common.cuh"
#include <math.h>

namespace evfly {

// numpy's floor division for doubles (numpy/_core/src/npymath/npy_math_internal.h.src, npy_divmod)
__device__ __forceinline__ double npy_floor_divide(double a, double b) {
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0) {
        if ((b < 0.0) != (mod < 0.0)) div -= 1.0;
    }
    double floordiv;
    if (div != 0.0) {
        floordiv = floor(div);
        if (div - floordiv > 0.5) floordiv += 1.0;
    } else {
        floordiv = copysign(0.0, a / b);
    }
    return floordiv;
}

// pass 1: difflog = log(im + eps) - log(prev + eps) (or im - prev when the inputs are already logs),
// and the global max |difflog| (non-negative doubles order like their bit patterns)
__global__ void __launch_bounds__(256)
k_difflog_pass1(const double* __restrict__ im, const double* __restrict__ prev, long long n, double eps, int inputs_are_log,
                double* __restrict__ difflog, unsigned long long* __restrict__ absmax_bits) {
    double local = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double d = inputs_are_log ? __dsub_rn(im[i], prev[i]) : __dsub_rn(log(__dadd_rn(im[i], eps)), log(__dadd_rn(prev[i], eps)));
        difflog[i] = d;
        const double a = fabs(d);
        if (a > local || a != a) local = a;   // NaN propagates to the max like np.abs(x).max()
    }
    unsigned long long bits = (unsigned long long)__double_as_longlong(local);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = other > bits ? other : bits;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(absmax_bits, bits);
}

// pass 2: events[d > 0] = (d // pos) * pos ; events[d < 0] = (d // -neg) * -neg ; zeros if max|d| < max(pos, neg)
__global__ void __launch_bounds__(256)
k_difflog_pass2(double* __restrict__ difflog_inout, long long n, double pos, double neg, const unsigned long long* __restrict__ absmax_bits) {
    const double amax = __longlong_as_double((long long)*absmax_bits);
    const bool all_zero = amax < fmax(pos, neg);      // false when amax is NaN, like the reference's `if ... < ...: return`
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double d = difflog_inout[i];
        double e = 0.0;
        if (!all_zero) {
            if (d > 0.0) e = __dmul_rn(npy_floor_divide(d, pos), pos);
            else if (d < 0.0) e = __dmul_rn(npy_floor_divide(d, -neg), -neg);
        }
        difflog_inout[i] = e;
    }
}

__global__ void __launch_bounds__(256) k_min_cutoff(float* __restrict__ x, long long n, float cutoff) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (fabsf(x[i]) < cutoff) x[i] = 0.f;
}

}  // namespace evfly

using namespace evfly;

extern "C" int evfly_difflog_events_f64(const double* d_im, const double* d_prev, int64_t n, double eps, int inputs_are_log,
                                        double pos_thresh, double neg_thresh, double* d_events, void* d_ws8, void* stream) {
    EVFLY_REQUIRE(d_im && d_prev && d_events && d_ws8 && n >= 0 && pos_thresh > 0 && neg_thresh > 0, "difflog_events_f64: bad argument");
    if (n == 0) return EVFLY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    EVFLY_CUDA(cudaMemsetAsync(d_ws8, 0, 8, st));
    const int grid = stream_grid(n, 256 * 4, 8);
    k_difflog_pass1<<<grid, 256, 0, st>>>(d_im, d_prev, n, eps, inputs_are_log, d_events, reinterpret_cast<unsigned long long*>(d_ws8));
    EVFLY_LAUNCHED();
    k_difflog_pass2<<<grid, 256, 0, st>>>(d_events, n, pos_thresh, neg_thresh, reinterpret_cast<const unsigned long long*>(d_ws8));
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_min_cutoff_f32(float* d_x, int64_t n, float cutoff, void* stream) {
    EVFLY_REQUIRE(d_x && n >= 0, "min_cutoff_f32: bad argument");
    if (n == 0) return EVFLY_OK;
    k_min_cutoff<<<stream_grid(n, 256 * 4, 16), 256, 0, (cudaStream_t)stream>>>(d_x, n, cutoff);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}
