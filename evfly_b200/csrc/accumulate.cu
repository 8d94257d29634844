// accumulate.cu -- L1 of the hot path: packed event records -> count frames, u8 node frames,
// temporal-bilinear voxel grids, single window or T windows of one stream.
//
// Replaces (reference paths relative to the evfly tree):
//   evfly_ros/src/node.cpp:29-39       per-event ++/-- on a 128-biased u8 image (wraps)
//   evfly_dv_ros/src/node.cpp:29-44    same, saturating at 0/255
//   utils/ev_utils.py:137-141,155-159  two boolean-mask copies + two np.histogram2d
//   utils/to_events.py:400-411         T rescans of the whole stream, one per window
//
// Data layout in HBM: events are an array of 16-byte evfly_event records (one coalesced
// ld.global.v4 per event, 512 B per warp); outputs are planar int32 [2,H,W] (plane 0 = negative
// counts, plane 1 = positive counts) and fp32 [B,H,W]. The whole output of one window
// (8.6 MB at 480x640, B=5) lives in the 126 MB L2 while events stream through, so the
// scattered updates are L2 reductions (RED), never DRAM read-modify-writes.
//
// All kernels are HBM-streaming integer/byte work: no tensor cores (deliberately).
#include "common.cuh"

namespace evfly {

struct Ev {
    unsigned x, y, pol;
    int64_t t_ns;
};

__device__ __forceinline__ Ev decode(const uint4& r) {
    Ev e;
    e.x = r.x & 0xffffu;
    e.y = r.x >> 16;
    e.pol = r.w & 0xffu;
    e.t_ns = (int64_t)r.y * 1000000000ll + (int64_t)r.z;
    return e;
}

__device__ __forceinline__ uint4 skip_record() { return make_uint4(0u, 0u, 0u, EVFLY_POL_SKIP); }

// ---------------------------------------------------------------------------------------
// count frames: counts[pol][y][x] += 1
// ---------------------------------------------------------------------------------------
template <int UNROLL>
__global__ void __launch_bounds__(256)
k_accumulate_counts(const uint4* __restrict__ ev, int64_t n, unsigned H, unsigned W,
                    int* __restrict__ counts) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const unsigned HW = H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * UNROLL) {
        uint4 r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t j = i + u * stride;
            r[u] = (j < n) ? ld_stream_v4(ev + j) : skip_record();
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const Ev e = decode(r[u]);
            if (e.x < W && e.y < H && e.pol < 2u)
                atomicAdd(counts + e.pol * HW + e.y * W + e.x, 1);  // result unused -> RED.ADD
        }
    }
}

// ---------------------------------------------------------------------------------------
// one window, direct: RED.s32 into counts + two RED.f32 into the voxel planes
// ---------------------------------------------------------------------------------------
struct VoxelParams {
    int64_t t0_ns, t1_ns;
    double scale;  // (B-1) / (t1 - t0)
    int B;
};

__device__ __forceinline__ void voxel_weights(const VoxelParams& p, int64_t t_ns, int& seg,
                                              float& w_lo, float& w_hi) {
    // tau in [0, B-1): computed in fp64 so that the fp32 weights are correctly rounded
    const double tau = (double)(t_ns - p.t0_ns) * p.scale;
    int s = (int)tau;
    if (s > p.B - 2) s = p.B - 2;
    if (s < 0) s = 0;
    const double f = tau - (double)s;
    seg = s;
    w_hi = (float)f;
    w_lo = (float)(1.0 - f);
}

template <int UNROLL>
__global__ void __launch_bounds__(256)
k_voxel_direct(const uint4* __restrict__ ev, int64_t n, unsigned H, unsigned W, VoxelParams p,
               int* __restrict__ counts, float* __restrict__ voxel) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const unsigned HW = H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * UNROLL) {
        uint4 r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t j = i + u * stride;
            r[u] = (j < n) ? ld_stream_v4(ev + j) : skip_record();
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const Ev e = decode(r[u]);
            if (e.x < W && e.y < H && e.pol < 2u && e.t_ns >= p.t0_ns && e.t_ns < p.t1_ns) {
                const unsigned pix = e.y * W + e.x;
                if (counts) atomicAdd(counts + e.pol * HW + pix, 1);
                if (voxel) {
                    const float sgn = e.pol ? 1.0f : -1.0f;
                    if (p.B == 1) {
                        atomicAdd(voxel + pix, sgn);
                    } else {
                        int s;
                        float w_lo, w_hi;
                        voxel_weights(p, e.t_ns, s, w_lo, w_hi);
                        atomicAdd(voxel + (size_t)s * HW + pix, sgn * w_lo);
                        atomicAdd(voxel + (size_t)(s + 1) * HW + pix, sgn * w_hi);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// one window, staged: ONE 16-byte vector reduction per event into an L2-resident stage
// stage[seg][y][x] = float4{ #pos, #neg, sum_pos f, sum_neg f }  (f = tau - seg in [0,1))
// then a finalise kernel turns the (B-1) segment planes into counts + B voxel planes and
// re-zeroes the stage. Counts are exact integers in fp32 up to 2^24 per (pixel,polarity,seg).
// ---------------------------------------------------------------------------------------
template <int UNROLL>
__global__ void __launch_bounds__(256)
k_voxel_staged(const uint4* __restrict__ ev, int64_t n, unsigned H, unsigned W, VoxelParams p,
               float4* __restrict__ stage) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const unsigned HW = H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * UNROLL) {
        uint4 r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t j = i + u * stride;
            r[u] = (j < n) ? ld_stream_v4(ev + j) : skip_record();
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const Ev e = decode(r[u]);
            if (e.x < W && e.y < H && e.pol < 2u && e.t_ns >= p.t0_ns && e.t_ns < p.t1_ns) {
                int s = 0;
                float w_lo = 1.0f, w_hi = 0.0f;
                if (p.B > 1) voxel_weights(p, e.t_ns, s, w_lo, w_hi);
                const float4 v = e.pol ? make_float4(1.0f, 0.0f, w_hi, 0.0f)
                                       : make_float4(0.0f, 1.0f, 0.0f, w_hi);
                atomicAdd(stage + (size_t)s * HW + e.y * W + e.x, v);  // REDG.E.ADD.F32x4
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_voxel_finalize(float4* __restrict__ stage, unsigned HW, int B, int* __restrict__ counts,
                 float* __restrict__ voxel) {
    const unsigned pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const int nseg = B > 1 ? B - 1 : 1;
    float npos = 0.f, nneg = 0.f;
    float carry = 0.f;  // (sum_pos f - sum_neg f) of the previous segment
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < nseg; ++s) {
        float4* q = stage + (size_t)s * HW + pix;
        const float4 v = *q;
        *q = zero;
        npos += v.x;
        nneg += v.y;
        const float c = v.x - v.y;  // exact: integers
        const float f = v.z - v.w;
        if (voxel) voxel[(size_t)s * HW + pix] = (c - f) + carry;
        carry = f;
    }
    if (voxel && B > 1) voxel[(size_t)(B - 1) * HW + pix] = carry;
    if (counts) {
        counts[pix] = (int)nneg;
        counts[HW + pix] = (int)npos;
    }
}

// ---------------------------------------------------------------------------------------
// T windows of one stream
// ---------------------------------------------------------------------------------------
// r[w] = first event index with t >= edges[w] (stream sorted by time)
// One warp per edge, 33-ary search: every round the 32 lanes probe 32 evenly spaced records of the current interval
// (5 dependent rounds for 10 M events instead of 24 for a scalar binary search).
__global__ void __launch_bounds__(128)
k_window_ranges(const uint4* __restrict__ ev, int64_t n, const int64_t* __restrict__ edges, int T, int64_t* __restrict__ ranges) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w > T) return;
    const int64_t key = edges[w];
    int64_t lo = 0, hi = n;          // invariant: t[i] < key for i < lo, t[i] >= key for i >= hi
    while (lo < hi) {
        const int64_t len = hi - lo;
        const int64_t step = (len + 32) / 33;                       // >= 1
        const int64_t pos = lo + (int64_t)(lane + 1) * step - 1;     // probes lo+step-1, lo+2*step-1, ...
        bool ge = true;                                              // positions past the interval count as >= key
        if (pos < hi) {
            const uint4 r = ev[pos];
            ge = ((int64_t)r.y * 1000000000ll + (int64_t)r.z) >= key;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ge);
        if (m == 0) {                                                // all 32 probes lie inside and are < key
            lo += 32 * step;
            continue;
        }
        const int first = __ffs(m) - 1;                              // first probe with t >= key (or the first one past hi)
        const int64_t new_hi = lo + (int64_t)(first + 1) * step - 1; // that probe's position (or beyond hi)
        const int64_t new_lo = lo + (int64_t)first * step;           // one past the previous probe
        hi = new_hi < hi ? new_hi : hi;
        lo = new_lo;
    }
    if (lane == 0) ranges[w] = lo;
}

// events [ranges[w0], ranges[w1]) (or the whole stream when ranges == nullptr) scattered into
// windows w0..w1-1; every event finds its window by binary search over the edges in smem.
template <int UNROLL>
__global__ void __launch_bounds__(256)
k_scatter_windows(const uint4* __restrict__ ev, int64_t n, const int64_t* __restrict__ edges,
                  int T, int w0, int w1, const int64_t* __restrict__ ranges, unsigned H,
                  unsigned W, int B, int* __restrict__ counts, float* __restrict__ voxel) {
    extern __shared__ int64_t s_edges[];  // edges[w0 .. w1]
    const int ne = w1 - w0 + 1;
    for (int k = threadIdx.x; k < ne; k += blockDim.x) s_edges[k] = edges[w0 + k];
    __syncthreads();
    int64_t begin = 0, end = n;
    if (ranges) {
        begin = ranges[w0];
        end = ranges[w1];
    }
    const int64_t t_first = s_edges[0], t_last = s_edges[ne - 1];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const unsigned HW = H * W;
    for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end;
         i += stride * UNROLL) {
        uint4 r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t j = i + u * stride;
            r[u] = (j < end) ? ld_stream_v4(ev + j) : skip_record();
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const Ev e = decode(r[u]);
            if (!(e.x < W && e.y < H && e.pol < 2u && e.t_ns >= t_first && e.t_ns < t_last))
                continue;
            // upper_bound(edges, t) - 1
            int lo = 0, hi = ne - 1;  // invariant: s_edges[lo] <= t < s_edges[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_edges[mid] <= e.t_ns) lo = mid; else hi = mid;
            }
            const int64_t ta = s_edges[lo], tb = s_edges[lo + 1];
            if (tb <= ta) continue;  // empty window
            const size_t w = (size_t)(w0 + lo);
            const unsigned pix = e.y * W + e.x;
            atomicAdd(counts + (w * 2 + e.pol) * HW + pix, 1);
            if (voxel) {
                const float sgn = e.pol ? 1.0f : -1.0f;
                float* v = voxel + w * B * HW;
                if (B == 1) {
                    atomicAdd(v + pix, sgn);
                } else {
                    VoxelParams p;
                    p.t0_ns = ta;
                    p.t1_ns = tb;
                    p.B = B;
                    p.scale = (double)(B - 1) / (double)(tb - ta);
                    int s;
                    float w_lo, w_hi;
                    voxel_weights(p, e.t_ns, s, w_lo, w_hi);
                    atomicAdd(v + (size_t)s * HW + pix, sgn * w_lo);
                    atomicAdd(v + (size_t)(s + 1) * HW + pix, sgn * w_hi);
                }
            }
        }
    }
}

// zero-fill of `n4` 32-bit words; 16-byte stores when the base is 16-byte aligned
__global__ void __launch_bounds__(256) k_zero_fill(uint32_t* __restrict__ p, int64_t n4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t done = 0;
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
        const int64_t n16 = n4 >> 2;
        uint4* q = reinterpret_cast<uint4*>(p);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int64_t i = tid; i < n16; i += stride) q[i] = z;
        done = n16 << 2;
    }
    for (int64_t i = done + tid; i < n4; i += stride) p[i] = 0u;
}

// ---------------------------------------------------------------------------------------
// counts -> float64 frame / u8 node frame
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_counts_to_frame_f64(const int* __restrict__ counts, unsigned HW, double pos_thresh,
                      double neg_thresh, double* __restrict__ frame) {
    const unsigned pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    // two rounded products and one rounded subtraction, exactly like
    // pos_thresh*hist_pos - neg_thresh*hist_neg in numpy (ev_utils.py:158); __dmul_rn/__dsub_rn
    // forbid the compiler from contracting this into an FMA.
    const double a = __dmul_rn(pos_thresh, (double)counts[HW + pix]);
    const double b = __dmul_rn(neg_thresh, (double)counts[pix]);
    frame[pix] = __dsub_rn(a, b);
}

__global__ void __launch_bounds__(256)
k_counts_to_u8(const int* __restrict__ counts, unsigned HW, int mode,
               const uint8_t* __restrict__ state_in, uint8_t* __restrict__ out,
               int* __restrict__ flagged, int flagged_cap, int* __restrict__ n_flagged) {
    const unsigned pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const int nneg = counts[pix], npos = counts[HW + pix];
    const int s = state_in ? (int)state_in[pix] : 128;
    int v = s + npos - nneg;
    if (mode == EVFLY_U8_WRAP) {
        out[pix] = (uint8_t)(v & 0xff);
    } else {
        if (s + npos > 255 || s - nneg < 0) {  // a clamp may have fired: order dependent
            const int slot = atomicAdd(n_flagged, 1);
            if (slot < flagged_cap) flagged[slot] = (int)pix;
        }
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        out[pix] = (uint8_t)v;
    }
}

// clamp-add maps v -> min(max(v + a, lo), hi) are closed under composition, so the node's
// sequential "if (v < 255) v++ / if (v > 0) v--" walk is an (associative, non-commutative)
// monoid fold over the events that hit one pixel, in stream order.
struct ClampAdd {
    int a, lo, hi;
};
__device__ __forceinline__ ClampAdd ca_identity() { return ClampAdd{0, -(1 << 30), 1 << 30}; }
__device__ __forceinline__ ClampAdd ca_then(const ClampAdd& f, const ClampAdd& g) {  // g after f
    ClampAdd r;
    r.a = f.a + g.a;
    r.hi = min(max(f.hi + g.a, g.lo), g.hi);
    r.lo = min(max(f.lo + g.a, g.lo), r.hi);
    return r;
}

__global__ void __launch_bounds__(256)
k_u8_saturate_replay(const uint4* __restrict__ ev, int64_t n, unsigned H, unsigned W,
                     const uint8_t* __restrict__ state_in, const int* __restrict__ flagged,
                     uint8_t* __restrict__ out) {
    __shared__ ClampAdd s_part[256];
    const unsigned pix = (unsigned)flagged[blockIdx.x];
    const unsigned px = pix % W, py = pix / W;
    const int64_t chunk = (n + blockDim.x - 1) / blockDim.x;
    const int64_t b = (int64_t)threadIdx.x * chunk;
    const int64_t e_end = min(n, b + chunk);
    ClampAdd f = ca_identity();
    const ClampAdd inc = ClampAdd{1, 0, 255}, dec = ClampAdd{-1, 0, 255};
    for (int64_t i = b; i < e_end; ++i) {
        const uint4 r = ev[i];
        const unsigned x = r.x & 0xffffu, y = r.x >> 16, pol = r.w & 0xffu;
        if (x == px && y == py && pol < 2u) f = ca_then(f, pol ? inc : dec);
    }
    s_part[threadIdx.x] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
        ClampAdd t = s_part[0];
        for (int k = 1; k < (int)blockDim.x; ++k) t = ca_then(t, s_part[k]);
        const int s = state_in ? (int)state_in[pix] : 128;
        out[pix] = (uint8_t)min(max(s + t.a, t.lo), t.hi);
    }
}

// used by accumulate_tiled.cu (kernels cannot be launched across translation units without -rdc)
int launch_window_ranges(const void* ev, int64_t n, const int64_t* edges, int T, int64_t* ranges, cudaStream_t st) {
    k_window_ranges<<<(T + 1 + 3) / 4, 128, 0, st>>>(reinterpret_cast<const uint4*>(ev), n, edges, T, ranges);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

}  // namespace evfly

// =========================================================================================
// C ABI
// =========================================================================================
using namespace evfly;

static inline bool dims_ok(int H, int W) { return H > 0 && W > 0 && H <= 65535 && W <= 65535 && (int64_t)H * W < (1ll << 30); }

extern "C" int evfly_accumulate_counts(const evfly_event* d_events, int64_t n, int H, int W,
                                       int32_t* d_counts, void* stream) {
    EVFLY_REQUIRE(n >= 0 && dims_ok(H, W), "accumulate_counts: bad n=%lld H=%d W=%d", (long long)n, H, W);
    EVFLY_REQUIRE(d_counts && (d_events || n == 0), "accumulate_counts: null pointer");
    if (n == 0) return EVFLY_OK;
    constexpr int U = 4;
    const int grid = stream_grid(n, 256 * U, 8);
    k_accumulate_counts<U><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(d_events), n, (unsigned)H, (unsigned)W, d_counts);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_counts_to_frame_f64(const int32_t* d_counts, int H, int W, double pos_thresh,
                                         double neg_thresh, double* d_frame, void* stream) {
    EVFLY_REQUIRE(dims_ok(H, W) && d_counts && d_frame, "counts_to_frame_f64: bad argument");
    const unsigned HW = (unsigned)H * W;
    k_counts_to_frame_f64<<<(HW + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        d_counts, HW, pos_thresh, neg_thresh, d_frame);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_counts_to_u8(const int32_t* d_counts, int H, int W, int mode,
                                  const uint8_t* d_state_in, uint8_t* d_frame_out,
                                  int32_t* d_flagged, int32_t flagged_cap, int32_t* d_n_flagged,
                                  void* stream) {
    EVFLY_REQUIRE(dims_ok(H, W) && d_counts && d_frame_out, "counts_to_u8: bad argument");
    EVFLY_REQUIRE(mode == EVFLY_U8_WRAP || mode == EVFLY_U8_SATURATE, "counts_to_u8: bad mode %d", mode);
    EVFLY_REQUIRE(mode == EVFLY_U8_WRAP || (d_flagged && d_n_flagged && flagged_cap > 0),
                  "counts_to_u8: SATURATE needs d_flagged / d_n_flagged");
    const unsigned HW = (unsigned)H * W;
    k_counts_to_u8<<<(HW + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        d_counts, HW, mode, d_state_in, d_frame_out, d_flagged, flagged_cap, d_n_flagged);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_u8_saturate_replay(const evfly_event* d_events, int64_t n, int H, int W,
                                        const uint8_t* d_state_in, const int32_t* d_flagged,
                                        int32_t n_flagged, uint8_t* d_frame_out, void* stream) {
    EVFLY_REQUIRE(n >= 0 && dims_ok(H, W) && d_frame_out && n_flagged >= 0, "u8_saturate_replay: bad argument");
    if (n_flagged == 0) return EVFLY_OK;
    EVFLY_REQUIRE(d_events && d_flagged, "u8_saturate_replay: null pointer");
    k_u8_saturate_replay<<<n_flagged, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(d_events), n, (unsigned)H, (unsigned)W, d_state_in,
        d_flagged, d_frame_out);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int64_t evfly_voxel_workspace_bytes(int H, int W, int B) {
    if (!dims_ok(H, W) || B < 1) return 0;
    return (int64_t)H * W * 16 * (B > 1 ? B - 1 : 1);
}

extern "C" int evfly_voxelize_window(const evfly_event* d_events, int64_t n, int H, int W, int B,
                                     int64_t t0_ns, int64_t t1_ns, int32_t* d_counts,
                                     float* d_voxel, void* d_ws, int algo, void* stream) {
    EVFLY_REQUIRE(n >= 0 && dims_ok(H, W) && B >= 1 && B <= 64, "voxelize_window: bad n/H/W/B");
    EVFLY_REQUIRE(t1_ns > t0_ns, "voxelize_window: empty window [%lld,%lld)", (long long)t0_ns, (long long)t1_ns);
    EVFLY_REQUIRE(d_counts || d_voxel, "voxelize_window: no output requested");
    EVFLY_REQUIRE(d_events || n == 0, "voxelize_window: null events");
    EVFLY_REQUIRE(algo == 0 || (algo == 1 && d_ws), "voxelize_window: algo %d needs a workspace", algo);
    VoxelParams p;
    p.t0_ns = t0_ns;
    p.t1_ns = t1_ns;
    p.B = B;
    p.scale = (double)(B - 1) / (double)(t1_ns - t0_ns);
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int U = 4;
    const unsigned HW = (unsigned)H * W;
    if (algo == 0) {
        if (n == 0) return EVFLY_OK;
        const int grid = stream_grid(n, 256 * U, 8);
        k_voxel_direct<U><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(d_events), n,
                                                (unsigned)H, (unsigned)W, p, d_counts, d_voxel);
        EVFLY_LAUNCHED();
    } else {
        if (n > 0) {
            const int grid = stream_grid(n, 256 * U, 8);
            k_voxel_staged<U><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(d_events), n,
                                                    (unsigned)H, (unsigned)W, p,
                                                    reinterpret_cast<float4*>(d_ws));
            EVFLY_LAUNCHED();
        }
        k_voxel_finalize<<<(HW + 255) / 256, 256, 0, st>>>(reinterpret_cast<float4*>(d_ws), HW, B,
                                                           d_counts, d_voxel);
        EVFLY_LAUNCHED();
    }
    return EVFLY_OK;
}

extern "C" int evfly_accumulate_windows(const evfly_event* d_events, int64_t n,
                                        const int64_t* d_edges_ns, int T, int H, int W, int B,
                                        int32_t* d_counts, float* d_voxel, int sorted_by_time,
                                        int64_t* d_range_ws, void* stream) {
    EVFLY_REQUIRE(n >= 0 && dims_ok(H, W) && T >= 1, "accumulate_windows: bad n/H/W/T");
    EVFLY_REQUIRE(d_counts && d_edges_ns && (d_events || n == 0), "accumulate_windows: null pointer");
    EVFLY_REQUIRE(!d_voxel || (B >= 1 && B <= 64), "accumulate_windows: bad B=%d", B);
    EVFLY_REQUIRE(!sorted_by_time || d_range_ws, "accumulate_windows: sorted mode needs d_range_ws");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t HW = (int64_t)H * W;
    const int64_t cnt_bytes_w = HW * 2 * 4;
    const int64_t vox_bytes_w = d_voxel ? HW * B * 4 : 0;
    constexpr int U = 4;
    const uint4* ev = reinterpret_cast<const uint4*>(d_events);

    // group size: keep one group's outputs (zero-filled, then RED-updated) inside L2
    const int64_t l2_budget = 40ll << 20;
    int G = (int)(l2_budget / (cnt_bytes_w + vox_bytes_w));
    if (G < 1) G = 1;
    if (!sorted_by_time || n == 0) G = T;  // one pass over everything
    // edges of one group must fit the kernel's dynamic smem (48 KB without opt-in)
    const int max_edges = (48 * 1024) / 8;
    if (G + 1 > max_edges) G = max_edges - 1;
    EVFLY_REQUIRE(sorted_by_time || T + 1 <= max_edges, "accumulate_windows: unsorted streams support T <= %d", max_edges - 1);

    if (sorted_by_time && n > 0) {
        k_window_ranges<<<(T + 1 + 3) / 4, 128, 0, st>>>(ev, n, d_edges_ns, T, d_range_ws);
        EVFLY_LAUNCHED();
    }
    for (int w0 = 0; w0 < T; w0 += G) {
        const int w1 = (w0 + G < T) ? w0 + G : T;
        // zero-fill this group's outputs; they stay L2-resident for the scatter that follows
        {
            uint32_t* c0 = reinterpret_cast<uint32_t*>(d_counts) + (int64_t)w0 * (cnt_bytes_w / 4);
            const int64_t c4 = (int64_t)(w1 - w0) * (cnt_bytes_w / 4);
            k_zero_fill<<<stream_grid(c4, 256 * 16, 8), 256, 0, st>>>(c0, c4);
            EVFLY_LAUNCHED();
            if (d_voxel) {
                uint32_t* v0 = reinterpret_cast<uint32_t*>(d_voxel) + (int64_t)w0 * (vox_bytes_w / 4);
                const int64_t v4 = (int64_t)(w1 - w0) * (vox_bytes_w / 4);
                k_zero_fill<<<stream_grid(v4, 256 * 16, 8), 256, 0, st>>>(v0, v4);
                EVFLY_LAUNCHED();
            }
        }
        if (n == 0) continue;
        const int64_t est = sorted_by_time ? ceil_div(n * (w1 - w0), T) * 2 : n;
        const int grid = stream_grid(est, 256 * U, 8);
        const size_t smem = (size_t)(w1 - w0 + 1) * sizeof(int64_t);
        k_scatter_windows<U><<<grid, 256, smem, st>>>(ev, n, d_edges_ns, T, w0, w1,
                                                      sorted_by_time ? d_range_ws : nullptr,
                                                      (unsigned)H, (unsigned)W, B, d_counts, d_voxel);
        EVFLY_LAUNCHED();
    }
    return EVFLY_OK;
}
