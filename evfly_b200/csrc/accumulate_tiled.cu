// accumulate_tiled.cu -- windows of a stream whose events are already grouped by window (a time-sorted stream
// plus the index range of every window, or the 8-byte wire format below): count frames + temporal-bilinear voxel
// grids through SHARED-MEMORY histogram tiles, without a single global reduction.
//
// Replaces utils/to_events.py:400-411 (T rescans of the stream, two np.histogram2d per window) and, for one
// window, utils/ev_utils.py:150-161 / evfly_ros/src/node.cpp:29-39 -- like the scatter kernels of accumulate.cu,
// by the route BASELINE.json north_star names ("shared-memory histogram tiles"). The scatter kernels are bound by
// the L2 reduction rate (~135 G RED/s: 74 us for 10 M events, plus a zero-fill of the output); here HBM sees
// every event once and every output byte once:
//
//   pass 0  k_chunk_plan      one CTA: windows -> chunks of <= 8192 events (a chunk never straddles a window)
//   pass 1  k_chunk_sort      one CTA per chunk: events -> registers (coalesced 8/16-byte loads, all in flight at
//                             once), masks applied, band = row / rows_per_band ranked with ONE shared-memory
//                             atomic per event, counting sort in shared memory, the chunk written back as compact
//                             8-byte records ordered by band + the chunk's band offsets. No global atomics, no
//                             inter-CTA dependency.
//   pass 2  k_band_accumulate one CTA per (band, window): the band's 2 + B planes live in shared memory; the CTA
//                             walks the band's piece of every chunk of its window (each event read once, from L2
//                             when pass 1 has just written it), shared-memory atomics, coalesced write-out.
//                             Nothing is zero-filled in HBM and counts are exact integers whatever the order.
//
// Measured dead ends (B200, 10 M events @480x640; profiles/r2_accumulate_variants.txt): ranking pass 1 with match.any
// and per-warp counter tables instead of one shared-memory atomic per event: 77 -> 150 us (MATCH.ANY is slow and the
// 16 batches of a warp serialise); accumulating the voxel bins as 2^-24 fixed point with the native 32-bit integer
// atomic + a carry word instead of atomicAdd(float) (a CAS spin loop in SASS, sm_100 has no shared-memory fp32 or
// 64-bit add): pass 2 88 -> 113 us (one window) and 270 -> 483 us (100 windows: 48 instead of 28 bytes of shared
// memory per pixel mean more, smaller bands). Both were reverted.
//
// Algorithmic bytes: record_bytes * events + 4*(2+B)*H*W per window. Extra traffic: 8 B per event written and read
// once more (the sorted copy; L2-resident for up to ~10 M events).
//
// 8-byte wire record (evfly_event8): the canonical 16-byte record with the absolute timestamp replaced by the
// offset from the first edge of the window the event belongs to: {u16 x, u16 y, u32 (dt_ns << 1) | polarity}.
// It halves the bytes per event over PCIe (the end-to-end bound of the offline evaluation path).
#include "common.cuh"

namespace evfly {

constexpr int kChunk = 8192;          // events per chunk
constexpr int kSortThreads = 512;     // 16 events per thread, all loads in flight before the first use
constexpr int kEPT = kChunk / kSortThreads;
constexpr int kMaxBands = 511;      // + the total = 512 scan elements, one per thread of pass 1
constexpr int kSeg = 8;               // pass 2: events handled sequentially by one lane

struct BandGeom {
    int rows, bands;
    uint32_t magic;      // band = umulhi(y, magic) for rows > 1
    size_t smem;         // pass-2 dynamic shared memory
};

static inline bool band_geometry(int H, int W, int planes, int n_windows, BandGeom* g) {
    const long long per_row = 4ll * planes * W;
    if (per_row > 200 * 1024) return false;
    long long rows_cap = (100 * 1024) / per_row;            // two CTAs per SM when the band allows it
    if (rows_cap < 1) rows_cap = 1;
    // few windows: cut the frame into enough bands to give every SM one
    long long want_bands = n_windows >= kNumSMs ? 1 : (kNumSMs + n_windows - 1) / n_windows;
    long long rows_par = (H + want_bands - 1) / want_bands;
    long long rows = rows_par < rows_cap ? rows_par : rows_cap;
    if (rows < 1) rows = 1;
    long long bands = (H + rows - 1) / rows;
    if (bands > kMaxBands) {
        rows = (H + kMaxBands - 1) / kMaxBands;
        bands = (H + rows - 1) / rows;
        if (rows * per_row > 200 * 1024) return false;
    }
    g->rows = (int)rows;
    g->bands = (int)bands;
    g->magic = rows > 1 ? (uint32_t)((1ull << 32) / (unsigned long long)rows + 1ull) : 0u;
    g->smem = (size_t)(rows * per_row) + 16;
    return true;
}

__device__ __forceinline__ uint2 ld_stream_v2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// ---- pass 0 -------------------------------------------------------------------------------
// chunk_first[w] = number of chunks of the windows before w; chunk_first[n_windows] = total
__global__ void __launch_bounds__(1024)
k_chunk_plan(const int64_t* __restrict__ win_offsets, int n_windows, int* __restrict__ chunk_first) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_windows; base += 1024) {
        const int w = base + threadIdx.x;
        int c = 0;
        if (w < n_windows) {
            const int64_t len = win_offsets[w + 1] - win_offsets[w];
            c = len > 0 ? (int)((len + kChunk - 1) / kChunk) : 0;
        }
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int v = s_warp[lane], iv = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, iv, d);
                if (lane >= d) iv += t;
            }
            s_warp[lane] = iv - v;      // exclusive
        }
        __syncthreads();
        const int carry = s_carry;
        if (w < n_windows) chunk_first[w] = carry + s_warp[warp] + incl - c;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_first[n_windows] = s_carry;
}

// ---- pass 1 -------------------------------------------------------------------------------
// sorted record: word0 = x | (row inside the band) << 16 | polarity << 31 ; word1 = t - t0 of the window (u32)
// REC = bytes per input record: 16 (evfly_event), 8 (evfly_event8) or 4 (evfly_event4: time as a 12-bit delta in microseconds to the
// previous record of the window; chunk_base_us[c] = offset of the last record before chunk c from the window's first edge)
template <int REC>
__global__ void __launch_bounds__(kSortThreads, 2)
k_chunk_sort(const void* __restrict__ records, const int64_t* __restrict__ win_offsets, const int64_t* __restrict__ win_t0,
             const int64_t* __restrict__ win_t1, int n_windows, unsigned H, unsigned W, unsigned rows, uint32_t magic, int bands,
             const int* __restrict__ chunk_first, int max_chunks, uint2* __restrict__ sorted, int* __restrict__ table_t,
             const uint32_t* __restrict__ chunk_base_us) {
    extern __shared__ __align__(16) uint2 s_sorted[];     // [kChunk]
    __shared__ int s_hist[kMaxBands + 1];
    __shared__ int s_warp[kSortThreads / 32];
    __shared__ int s_w;
    const int c = blockIdx.x;
    if (c >= chunk_first[n_windows]) return;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_windows;         // chunk_first[lo] <= c < chunk_first[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (chunk_first[mid] <= c) lo = mid; else hi = mid;
        }
        s_w = lo;
    }
    for (int b = threadIdx.x; b <= bands; b += kSortThreads) s_hist[b] = 0;
    __syncthreads();
    const int w = s_w;
    const int64_t ebeg = win_offsets[w] + (int64_t)(c - chunk_first[w]) * kChunk;
    const int64_t wend = win_offsets[w + 1];
    const int n_here = (int)(wend - ebeg < kChunk ? wend - ebeg : kChunk);
    const int64_t t0 = win_t0[w], t1 = win_t1[w];
    const bool live = t1 > t0;
    const uint64_t len = live ? (uint64_t)(t1 - t0) : 0ull;

    uint32_t w0[kEPT], w1[kEPT], key[kEPT];     // key = band << 16 | rank (rank < 8192), 0xffffffff = dropped
    if (REC == 16) {
        uint4 r[kEPT];
#pragma unroll
        for (int u = 0; u < kEPT; ++u) {
            const int k = u * kSortThreads + threadIdx.x;
            r[u] = k < n_here ? ld_stream_v4(reinterpret_cast<const uint4*>(records) + ebeg + k) : make_uint4(0xffffffffu, 0u, 0u, 2u);
        }
#pragma unroll
        for (int u = 0; u < kEPT; ++u) {
            const unsigned x = r[u].x & 0xffffu, y = r[u].x >> 16, pol = r[u].w & 0xffu;
            const int64_t t = (int64_t)r[u].y * 1000000000ll + (int64_t)r[u].z;
            const uint64_t dt = (uint64_t)(t - t0);
            const bool ok = x < W && y < H && pol < 2u && t >= t0 && dt < len && dt < (1ull << 32);
            key[u] = ok ? y : 0xffffffffu;
            w0[u] = x | (pol << 31);
            w1[u] = (uint32_t)dt;
        }
    } else if (REC == 4) {
        // coalesced 4-byte loads, then the running time of the chunk: the deltas go through shared memory (the buffer of the sorted
        // records is still free) so that every thread sums 16 CONSECUTIVE records, one block scan over the thread totals adds the
        // rest, and the absolute offsets come back the same way
        uint32_t r[kEPT];
#pragma unroll
        for (int u = 0; u < kEPT; ++u) {
            const int k = u * kSortThreads + threadIdx.x;
            r[u] = k < n_here ? __ldg(reinterpret_cast<const uint32_t*>(records) + ebeg + k) : 0x000fffffu;       // skip record, delta 0
        }
        uint32_t* s_dt = reinterpret_cast<uint32_t*>(s_sorted);                // [kChunk]
#pragma unroll
        for (int u = 0; u < kEPT; ++u) s_dt[u * kSortThreads + threadIdx.x] = r[u] >> 20;
        __syncthreads();
        {
            uint32_t loc[kEPT];
            const uint4* src = reinterpret_cast<const uint4*>(s_dt + threadIdx.x * kEPT);
            uint32_t run = 0;
#pragma unroll
            for (int q = 0; q < kEPT / 4; ++q) {
                const uint4 v = src[q];
                run += v.x; loc[4 * q] = run; run += v.y; loc[4 * q + 1] = run; run += v.z; loc[4 * q + 2] = run; run += v.w; loc[4 * q + 3] = run;
            }
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            uint32_t incl = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (lane == 31) s_warp[warp] = (int)incl;
            __syncthreads();
            if (warp == 0) {
                const uint32_t x = lane < kSortThreads / 32 ? (uint32_t)s_warp[lane] : 0u;
                uint32_t ix = x;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, ix, d);
                    if (lane >= d) ix += t;
                }
                if (lane < kSortThreads / 32) s_warp[lane] = (int)(ix - x);
            }
            __syncthreads();
            const uint32_t before = chunk_base_us[c] + (uint32_t)s_warp[warp] + incl - run;
            uint4* dst = reinterpret_cast<uint4*>(s_dt + threadIdx.x * kEPT);
#pragma unroll
            for (int q = 0; q < kEPT / 4; ++q) dst[q] = make_uint4(before + loc[4 * q], before + loc[4 * q + 1], before + loc[4 * q + 2], before + loc[4 * q + 3]);
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kEPT; ++u) {
            const unsigned x = r[u] & 1023u, y = (r[u] >> 10) & 511u, pol = (r[u] >> 19) & 1u;
            const uint64_t dt = (uint64_t)s_dt[u * kSortThreads + threadIdx.x] * 1000ull;       // microseconds -> the nanoseconds of the other formats
            const bool ok = x < W && y < H && dt < len && dt < (1ull << 32);
            key[u] = ok ? y : 0xffffffffu;
            w0[u] = x | (pol << 31);
            w1[u] = (uint32_t)dt;
        }
        __syncthreads();      // s_dt is about to become s_sorted again (the ranking below does not touch it, the scatter does)
    } else {
        uint2 r[kEPT];
#pragma unroll
        for (int u = 0; u < kEPT; ++u) {
            const int k = u * kSortThreads + threadIdx.x;
            r[u] = k < n_here ? ld_stream_v2(reinterpret_cast<const uint2*>(records) + ebeg + k) : make_uint2(0xffffffffu, 0u);
        }
#pragma unroll
        for (int u = 0; u < kEPT; ++u) {
            const unsigned x = r[u].x & 0xffffu, y = r[u].x >> 16, pol = r[u].y & 1u;
            const uint64_t dt = (uint64_t)(r[u].y >> 1);
            const bool ok = x < W && y < H && dt < len;
            key[u] = ok ? y : 0xffffffffu;
            w0[u] = x | (pol << 31);
            w1[u] = (uint32_t)dt;
        }
    }
#pragma unroll
    for (int u = 0; u < kEPT; ++u) {
        if (key[u] != 0xffffffffu) {
            const unsigned y = key[u];
            const unsigned band = rows > 1 ? __umulhi(y, magic) : y;     // exact for y < 2^16 (magic = floor(2^32/rows) + 1)
            w0[u] |= (y - band * rows) << 16;
            const int rank = atomicAdd(&s_hist[band], 1);
            key[u] = (band << 16) | (unsigned)rank;
        }
    }
    __syncthreads();
    // exclusive scan of s_hist[0..bands] (bands <= 512 = one element per thread)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int v = (int)threadIdx.x <= bands ? s_hist[threadIdx.x] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int x = lane < kSortThreads / 32 ? s_warp[lane] : 0;
            int ix = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, ix, d);
                if (lane >= d) ix += t;
            }
            if (lane < kSortThreads / 32) s_warp[lane] = ix - x;
        }
        __syncthreads();
        if ((int)threadIdx.x <= bands) s_hist[threadIdx.x] = s_warp[warp] + incl - v;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kEPT; ++u)
        if (key[u] != 0xffffffffu) s_sorted[s_hist[key[u] >> 16] + (int)(key[u] & 0xffffu)] = make_uint2(w0[u], w1[u]);
    __syncthreads();
    const int total = s_hist[bands];
    uint2* out = sorted + ebeg;
    for (int k = threadIdx.x; k < total; k += kSortThreads) out[k] = s_sorted[k];
    for (int b = threadIdx.x; b <= bands; b += kSortThreads) table_t[(size_t)b * max_chunks + c] = s_hist[b];
}

// ---- pass 2 -------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 2)
k_band_accumulate(const uint2* __restrict__ sorted, const int64_t* __restrict__ win_offsets, const int64_t* __restrict__ win_t0,
                  const int64_t* __restrict__ win_t1, const int* __restrict__ out_slot, int slot_stride, int slot_offset, unsigned H,
                  unsigned W, int B, int has_voxel,
                  unsigned rows_per_band, const int* __restrict__ chunk_first, int max_chunks, const int* __restrict__ table_t,
                  int* __restrict__ counts, float* __restrict__ voxel) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ int s_seg[1025];                 // exclusive prefix of the segment counts of up to 1024 chunks
    __shared__ int s_lo[1024];
    __shared__ int s_warp[16];
    const int band = blockIdx.x, w = blockIdx.y;
    const unsigned y0 = (unsigned)band * rows_per_band;
    const unsigned rows = min(rows_per_band, H - y0);
    const unsigned npx = rows * W;
    const unsigned plane = rows_per_band * W;
    int* s_cnt = reinterpret_cast<int*>(s_raw);                     // [2][plane]
    float* s_vox = reinterpret_cast<float*>(s_raw) + 2 * plane;     // [B][plane]
    {
        const unsigned words = (2 + (has_voxel ? B : 0)) * plane;
        uint4* z = reinterpret_cast<uint4*>(s_raw);
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        for (unsigned i = threadIdx.x; i < (words + 3) / 4; i += blockDim.x) z[i] = zero;
    }
    const int64_t t0 = win_t0[w], t1 = win_t1[w];
    const double scale = (B > 1 && t1 > t0) ? (double)(B - 1) / (double)(t1 - t0) : 0.0;
    const int c0 = chunk_first[w], c1 = chunk_first[w + 1];
    const int64_t wbeg = win_offsets[w];
    const int* tlo = table_t + (size_t)band * max_chunks;
    const int* thi = table_t + (size_t)(band + 1) * max_chunks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int cb = c0; cb < c1; cb += 1024) {
        const int nc = min(1024, c1 - cb);
        __syncthreads();                        // zero-fill done / previous batch consumed
        // segments per chunk -> exclusive prefix (two chunks per thread)
        int seg[2], lo2[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = 2 * (int)threadIdx.x + q;
            seg[q] = lo2[q] = 0;
            if (k < nc) {
                lo2[q] = tlo[cb + k];
                seg[q] = (thi[cb + k] - lo2[q] + kSeg - 1) / kSeg;
            }
        }
        const int mine = seg[0] + seg[1];
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int x = lane < 16 ? s_warp[lane] : 0;
            int ix = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, ix, d);
                if (lane >= d) ix += t;
            }
            if (lane < 16) s_warp[lane] = ix - x;
        }
        __syncthreads();
        {
            const int excl = s_warp[warp] + incl - mine;
            const int k = 2 * (int)threadIdx.x;
            if (k < nc) { s_seg[k] = excl; s_lo[k] = lo2[0]; }
            if (k + 1 < nc) { s_seg[k + 1] = excl + seg[0]; s_lo[k + 1] = lo2[1]; }
            if (threadIdx.x == 511) s_seg[1024] = excl + mine;     // total (threads past nc carry zeros)
        }
        __syncthreads();
        const int total_seg = s_seg[1024];
        for (int sg = threadIdx.x; sg < total_seg; sg += blockDim.x) {
            int lo = 0, hi = nc;                // s_seg[lo] <= sg < s_seg[hi] (s_seg[nc] := total)
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_seg[mid] <= sg) lo = mid; else hi = mid;
            }
            const int k = lo;
            const int first = s_lo[k] + (sg - s_seg[k]) * kSeg;
            const int last = min(first + kSeg, thi[cb + k]);
            const uint2* src = sorted + wbeg + (int64_t)(cb + k - c0) * kChunk;
            uint2 e[kSeg];
#pragma unroll
            for (int j = 0; j < kSeg; ++j) e[j] = first + j < last ? src[first + j] : make_uint2(0u, 0u);
#pragma unroll
            for (int j = 0; j < kSeg; ++j) {
                if (first + j >= last) break;
                const unsigned x = e[j].x & 0xffffu, yy = (e[j].x >> 16) & 0x7fffu, pol = e[j].x >> 31;
                const unsigned pix = yy * W + x;
                atomicAdd(s_cnt + pol * plane + pix, 1);
                if (has_voxel) {
                    const float sgn = pol ? 1.0f : -1.0f;
                    if (B == 1) {
                        atomicAdd(s_vox + pix, sgn);
                    } else {
                        // identical arithmetic to voxel_weights() of accumulate.cu: tau in fp64, weights rounded once
                        const double tau = (double)e[j].y * scale;
                        int s = (int)tau;
                        if (s > B - 2) s = B - 2;
                        if (s < 0) s = 0;
                        const double f = tau - (double)s;
                        atomicAdd(s_vox + (unsigned)s * plane + pix, sgn * (float)(1.0 - f));
                        atomicAdd(s_vox + (unsigned)(s + 1) * plane + pix, sgn * (float)f);
                    }
                }
            }
        }
    }
    __syncthreads();
    const size_t HW = (size_t)H * W;
    const size_t band0 = (size_t)y0 * W;
    const size_t slot = out_slot ? (size_t)out_slot[w] : (size_t)w * slot_stride + slot_offset;
    int* cw = counts + slot * 2 * HW + band0;
    for (unsigned i = threadIdx.x; i < npx; i += blockDim.x) {
        cw[i] = s_cnt[i];
        cw[HW + i] = s_cnt[plane + i];
    }
    if (has_voxel) {
        float* vw = voxel + slot * B * HW + band0;
        for (int b = 0; b < B; ++b)
            for (unsigned i = threadIdx.x; i < npx; i += blockDim.x) vw[(size_t)b * HW + i] = s_vox[(unsigned)b * plane + i];
    }
}

// 16-byte records of a time-sorted stream -> 8-byte wire records, position by position (record i -> record i);
// window of event i = the range [win_offsets[w], win_offsets[w+1]) it lies in. Events outside every window, with
// polarity >= 2 or with an offset that does not fit 31 bits become skip records (x = y = 0xffff).
__global__ void __launch_bounds__(256)
k_pack_ev8(const uint4* __restrict__ ev, int64_t n, const int64_t* __restrict__ edges, const int64_t* __restrict__ offsets, int T,
           uint2* __restrict__ out) {
    extern __shared__ int64_t s_off[];      // offsets[0..T]
    for (int k = threadIdx.x; k <= T; k += blockDim.x) s_off[k] = offsets[k];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint4 r = ld_stream_v4(ev + i);
        uint2 o = make_uint2(0xffffffffu, 0u);
        if (i >= s_off[0] && i < s_off[T] && (r.w & 0xffu) < 2u) {
            int lo = 0, hi = T;                 // s_off[lo] <= i < s_off[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_off[mid] <= i) lo = mid; else hi = mid;
            }
            const int64_t t = (int64_t)r.y * 1000000000ll + (int64_t)r.z;
            const int64_t dt = t - edges[lo];
            if (dt >= 0 && dt < (1ll << 31) && t < edges[lo + 1]) o = make_uint2(r.x, ((unsigned)dt << 1) | (r.w & 1u));
        }
        out[i] = o;
    }
}

int launch_window_ranges(const void* ev, int64_t n, const int64_t* edges, int T, int64_t* ranges, cudaStream_t st);   // accumulate.cu

struct SortedWs {
    int* chunk_first;      // [n_windows + 1]
    int* table_t;          // [bands + 1][max_chunks]
    uint2* sorted;         // [n]
    int max_chunks;
};

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static int64_t sorted_ws_bytes(int64_t n, int n_windows, int bands) {
    const int64_t max_chunks = n / kChunk + n_windows + 1;
    return align_up(4ll * (n_windows + 1), 256) + align_up(4ll * (bands + 1) * max_chunks, 256) + align_up(8ll * n, 256) + 256;
}

static SortedWs carve_ws(void* ws, int64_t n, int n_windows, int bands) {
    SortedWs s;
    s.max_chunks = (int)(n / kChunk + n_windows + 1);
    uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    s.chunk_first = reinterpret_cast<int*>(p);
    p += align_up(4ll * (n_windows + 1), 256);
    s.table_t = reinterpret_cast<int*>(p);
    p += align_up(4ll * (bands + 1) * s.max_chunks, 256);
    s.sorted = reinterpret_cast<uint2*>(p);
    return s;
}

template <int REC>
static int run_sorted(const void* recs, int64_t n, const int64_t* offs, const int64_t* t0, const int64_t* t1, const int* out_slot, int slot_stride,
                      int slot_offset, int n_windows, int H, int W, int B, int32_t* counts, float* voxel, void* ws, int64_t ws_bytes, cudaStream_t st,
                      const uint32_t* chunk_base_us = nullptr) {
    BandGeom g;
    EVFLY_REQUIRE(band_geometry(H, W, 2 + (voxel ? B : 0), n_windows, &g), "accumulate_windows (tiled): a row of %d pixels x %d planes does not fit shared memory", W, 2 + B);
    EVFLY_REQUIRE(n <= (1ll << 40), "accumulate_windows (tiled): too many events");
    if (ws_bytes < sorted_ws_bytes(n, n_windows, g.bands)) {
        set_error("accumulate_windows (tiled): workspace of %lld bytes, %lld needed", (long long)ws_bytes, (long long)sorted_ws_bytes(n, n_windows, g.bands));
        return EVFLY_ERR_WORKSPACE;
    }
    const SortedWs s = carve_ws(ws, n, n_windows, g.bands);
    // per device: these attributes belong to the device's context, and setting them is cheap
    EVFLY_CUDA(cudaFuncSetAttribute(k_band_accumulate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    EVFLY_CUDA(cudaFuncSetAttribute(k_chunk_sort<REC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChunk * 8));
    for (int w0 = 0; w0 < n_windows; w0 += 32768) {          // gridDim.y limit and chunk_first's int range
        const int nw = n_windows - w0 < 32768 ? n_windows - w0 : 32768;
        k_chunk_plan<<<1, 1024, 0, st>>>(offs + w0, nw, s.chunk_first);
        EVFLY_LAUNCHED();
        if (n > 0) {
            k_chunk_sort<REC><<<s.max_chunks, kSortThreads, kChunk * 8, st>>>(recs, offs + w0, t0 + w0, t1 + w0, nw, (unsigned)H, (unsigned)W, (unsigned)g.rows,
                                                                     g.magic, g.bands, s.chunk_first, s.max_chunks, s.sorted, s.table_t, chunk_base_us);
            EVFLY_LAUNCHED();
        }
        k_band_accumulate<<<dim3(g.bands, nw), 512, g.smem, st>>>(s.sorted, offs + w0, t0 + w0, t1 + w0, out_slot ? out_slot + w0 : nullptr, slot_stride,
                                                                  slot_offset + w0 * slot_stride, (unsigned)H, (unsigned)W, B, voxel ? 1 : 0,
                                                                  (unsigned)g.rows, s.chunk_first, s.max_chunks, s.table_t, counts, voxel);
        EVFLY_LAUNCHED();
    }
    return EVFLY_OK;
}

}  // namespace evfly

using namespace evfly;

static inline bool dims_ok_t(int H, int W) { return H > 0 && W > 0 && H <= 32767 && W <= 65534 && (int64_t)H * W < (1ll << 30); }

extern "C" int64_t evfly_accumulate_sorted_workspace_bytes(int64_t n, int n_windows, int H, int W, int B) {
    // the band count only shrinks when the voxel planes are dropped or more windows are given: size for the largest
    BandGeom g;
    if (n < 0 || n_windows < 0 || !dims_ok_t(H, W) || B < 0 || !band_geometry(H, W, 2 + B, 1, &g)) return 0;
    return sorted_ws_bytes(n, n_windows, g.bands) + align_up(8ll * (n_windows + 1), 256) + 256;   // + the window ranges of the 16-byte entry point
}

extern "C" int evfly_accumulate_windows_sorted(const evfly_event* d_events, int64_t n, const int64_t* d_edges_ns, int T, int H, int W, int B,
                                               int32_t* d_counts, float* d_voxel, int slot_stride, int slot_offset, void* d_ws, int64_t ws_bytes,
                                               void* stream) {
    EVFLY_REQUIRE(n >= 0 && dims_ok_t(H, W) && T >= 1 && slot_stride >= 1 && slot_offset >= 0, "accumulate_windows_sorted: bad n/H/W/T/slot");
    EVFLY_REQUIRE(d_counts && d_edges_ns && d_ws && (d_events || n == 0), "accumulate_windows_sorted: null pointer");
    EVFLY_REQUIRE(!d_voxel || (B >= 1 && B <= 64), "accumulate_windows_sorted: bad B=%d", B);
    EVFLY_REQUIRE(ws_bytes >= 8ll * (T + 1) + 256, "accumulate_windows_sorted: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    // window ranges of the time-sorted stream go to the front of the workspace
    int64_t* ranges = reinterpret_cast<int64_t*>((reinterpret_cast<uintptr_t>(d_ws) + 255) & ~(uintptr_t)255);
    if (n > 0) {
        const int rc = launch_window_ranges(d_events, n, d_edges_ns, T, ranges, st);
        if (rc) return rc;
    } else {
        EVFLY_CUDA(cudaMemsetAsync(ranges, 0, sizeof(int64_t) * (T + 1), st));
    }
    const int64_t used = align_up(8ll * (T + 1), 256) + 256;
    return run_sorted<16>(d_events, n, ranges, d_edges_ns, d_edges_ns + 1, nullptr, slot_stride, slot_offset, T, H, W, B, d_counts, d_voxel,
                            reinterpret_cast<uint8_t*>(d_ws) + used, ws_bytes - used, st);
}

extern "C" int evfly_accumulate_windows_ev8(const evfly_event8* d_events, int64_t n, const int64_t* d_win_offsets, const int64_t* d_win_t0,
                                            const int64_t* d_win_t1, const int32_t* d_out_slot, int n_windows, int H, int W, int B,
                                            int32_t* d_counts, float* d_voxel, void* d_ws, int64_t ws_bytes, void* stream) {
    EVFLY_REQUIRE(n >= 0 && dims_ok_t(H, W) && n_windows >= 0, "accumulate_windows_ev8: bad n/H/W/n_windows");
    EVFLY_REQUIRE(d_counts && d_win_offsets && d_win_t0 && d_win_t1 && d_ws && (d_events || n == 0), "accumulate_windows_ev8: null pointer");
    EVFLY_REQUIRE(!d_voxel || (B >= 1 && B <= 64), "accumulate_windows_ev8: bad B=%d", B);
    EVFLY_REQUIRE((reinterpret_cast<uintptr_t>(d_events) & 7) == 0, "accumulate_windows_ev8: records must be 8-byte aligned");
    if (n_windows == 0) return EVFLY_OK;
    return run_sorted<8>(d_events, n, d_win_offsets, d_win_t0, d_win_t1, d_out_slot, 1, 0, n_windows, H, W, B, d_counts, d_voxel, d_ws, ws_bytes,
                             (cudaStream_t)stream);
}

extern "C" int evfly_accumulate_chunk_events(void) { return kChunk; }

extern "C" int evfly_accumulate_windows_ev4(const uint32_t* d_events, int64_t n, const int64_t* d_win_offsets, const int64_t* d_win_t0,
                                            const int64_t* d_win_t1, const int32_t* d_out_slot, const uint32_t* d_chunk_base_us, int n_windows, int H,
                                            int W, int B, int32_t* d_counts, float* d_voxel, void* d_ws, int64_t ws_bytes, void* stream) {
    EVFLY_REQUIRE(n >= 0 && H > 0 && W > 0 && H <= 511 && W <= 1023 && n_windows >= 0 && n_windows <= 32768,
                  "accumulate_windows_ev4: needs H <= 511, W <= 1023 (9 + 10 coordinate bits; (1023, 511) is the skip record) and at most 32768 windows");
    EVFLY_REQUIRE(d_counts && d_win_offsets && d_win_t0 && d_win_t1 && d_chunk_base_us && d_ws && (d_events || n == 0), "accumulate_windows_ev4: null pointer");
    EVFLY_REQUIRE(!d_voxel || (B >= 1 && B <= 64), "accumulate_windows_ev4: bad B=%d", B);
    EVFLY_REQUIRE((reinterpret_cast<uintptr_t>(d_events) & 3) == 0, "accumulate_windows_ev4: records must be 4-byte aligned");
    if (n_windows == 0) return EVFLY_OK;
    return run_sorted<4>(d_events, n, d_win_offsets, d_win_t0, d_win_t1, d_out_slot, 1, 0, n_windows, H, W, B, d_counts, d_voxel, d_ws, ws_bytes,
                         (cudaStream_t)stream, d_chunk_base_us);
}

extern "C" int evfly_pack_events_ev8(const evfly_event* d_events, int64_t n, const int64_t* d_edges_ns, int T, evfly_event8* d_out,
                                     int64_t* d_win_offsets, void* stream) {
    EVFLY_REQUIRE(n >= 0 && T >= 1 && T + 1 <= (48 * 1024) / 8, "pack_events_ev8: bad n / T (T <= 6143)");
    EVFLY_REQUIRE(d_edges_ns && d_win_offsets && ((d_events && d_out) || n == 0), "pack_events_ev8: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        EVFLY_CUDA(cudaMemsetAsync(d_win_offsets, 0, sizeof(int64_t) * (T + 1), st));
        return EVFLY_OK;
    }
    const int rc = launch_window_ranges(d_events, n, d_edges_ns, T, d_win_offsets, st);
    if (rc) return rc;
    k_pack_ev8<<<stream_grid(n, 256 * 4, 8), 256, sizeof(int64_t) * (T + 1), st>>>(reinterpret_cast<const uint4*>(d_events), n, d_edges_ns,
                                                                                  d_win_offsets, T, reinterpret_cast<uint2*>(d_out));
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}
