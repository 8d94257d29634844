// accumulate_binned.cu -- count frames for LARGE event batches by spatial binning + shared-memory histograms.
//
// k_accumulate_counts issues one L2 reduction per event, and the L2 reduction rate tops out at ~135 G/s on
// B200 whatever the launch shape (profiles/r1_microbench_atomics.json): 10 M events cannot go below ~74 us,
// 0.35 of the HBM roofline. Shared-memory integer atomics are 10x faster (1.44 T/s), but a 480x640x2 frame does
// not fit one SM, so events are first routed to the SM that owns their pixels:
//   K1 k_bin_events   : the frame is cut into tiles of 2048 consecutive pixels; a CTA takes 4096 events, ranks them
//                       per tile with shared-memory atomics, reserves space in each tile's global bucket with ONE
//                       global atomic per (CTA, tile), stages the 2-byte records {pixel-in-tile:11, polarity:1}
//                       sorted by tile in shared memory and writes them out in coalesced runs.
//                       A record that does not fit its bucket (pathological clustering) falls back to the L2 RED,
//                       so the result is exact for any distribution.
//   K2 k_bucket_counts: one CTA per tile streams its bucket (16-byte loads), accumulates into a 16 KB
//                       shared-memory tile with native integer atomics and adds the tile to the output; it owns
//                       its pixels, so the final update is a plain read-modify-write.
// Traffic: 16 B/event read + 2 B/event written + 2 B/event read + the frame: 1.25x the algorithmic bytes.
#include "common.cuh"

namespace evfly {

constexpr int kTilePix = 2048;      // pixels per tile (2 polarities x u32 = 16 KB of shared memory in K2)
constexpr int kTileShift = 11;
constexpr int kChunk = 4096;        // events per K1 iteration
constexpr int kEpt = 16;            // events per thread (256 threads)
constexpr int kMaxTiles = 4096;     // frames up to 8.4 M pixels
constexpr int kSub = 16;            // sub-buckets per tile: consecutive chunks reserve through different cursors, so the
                                    // chain of same-address L2 atomics (serialised, ~30 ns each) is 16x shorter
constexpr int kCursorBytes = kMaxTiles * kSub * 4;   // 256 KB header of the workspace

__global__ void __launch_bounds__(256, 6)
k_bin_events(const uint4* __restrict__ ev, long long n, unsigned H, unsigned W, int n_tiles, unsigned cap,
             unsigned* __restrict__ cursors, unsigned short* __restrict__ buckets, int* __restrict__ counts) {
    extern __shared__ unsigned s_mem[];
    unsigned* s_cnt = s_mem;                         // [n_tiles] events of this chunk per tile
    unsigned* s_off = s_cnt + n_tiles;               // [n_tiles] exclusive prefix inside the staging buffer
    unsigned* s_cur = s_off + n_tiles;               // [n_tiles] running insert position while staging
    unsigned* s_base = s_cur + n_tiles;              // [n_tiles] reserved start inside the global bucket
    unsigned* s_ev = s_base + n_tiles;               // [kChunk] compact event: pol << 30 | pixel, ~0 = dropped
    unsigned short* s_rec = reinterpret_cast<unsigned short*>(s_ev + kChunk);      // [kChunk] staged records
    unsigned short* s_tile = s_rec + kChunk;                                       // [kChunk] tile of each staged record
    __shared__ unsigned s_warp_tot[8];
    const unsigned HW = H * W;
    const long long n_chunks = (n + kChunk - 1) / kChunk;
    for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) s_cnt[t] = 0;
        __syncthreads();
        const long long base = chunk * kChunk;
        const unsigned sub = (unsigned)(chunk & (kSub - 1));
        // pass 1: stream the records (8 loads in flight per thread), keep 4 bytes per event, count per tile
#pragma unroll
        for (int u0 = 0; u0 < kEpt; u0 += 8) {
            unsigned xy[8], pw[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const long long i = base + (u0 + u) * 256 + threadIdx.x;
                xy[u] = 0xffffffffu;
                pw[u] = 0xffu;
                if (i < n) {
                    const uint4 r = ld_stream_v4(ev + i);
                    xy[u] = r.x;
                    pw[u] = r.w;
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned x = xy[u] & 0xffffu, y = xy[u] >> 16, pol = pw[u] & 0xffu;
                unsigned c = 0xffffffffu;
                if (x < W && y < H && pol < 2u) {
                    const unsigned pix = y * W + x;
                    c = pix | (pol << 30);
                    atomicAdd(&s_cnt[pix >> kTileShift], 1u);
                }
                s_ev[(u0 + u) * 256 + threadIdx.x] = c;
            }
        }
        __syncthreads();
        // exclusive scan of the per-tile counts (each thread scans a contiguous slice) and ONE global reservation
        // per non-empty tile
        {
            const int per = (n_tiles + 255) / 256;
            const int t0 = threadIdx.x * per, t1 = min(n_tiles, t0 + per);
            unsigned sum = 0;
            for (int t = t0; t < t1; ++t) sum += s_cnt[t];
            unsigned incl = sum;
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) s_warp_tot[wid] = incl;
            __syncthreads();
            unsigned woff = 0;
            for (int w = 0; w < wid; ++w) woff += s_warp_tot[w];
            unsigned run = woff + incl - sum;
            for (int t = t0; t < t1; ++t) {
                const unsigned c = s_cnt[t];
                s_off[t] = run;
                s_cur[t] = run;
                run += c;
                s_base[t] = c ? atomicAdd(&cursors[t * kSub + sub], c) : 0u;
            }
        }
        __syncthreads();
        // pass 2: stage the 2-byte records grouped by tile
#pragma unroll
        for (int u = 0; u < kEpt; ++u) {
            const unsigned c = s_ev[u * 256 + threadIdx.x];
            if (c != 0xffffffffu) {
                const unsigned pix = c & 0x3fffffffu, tile = pix >> kTileShift;
                const unsigned pos = atomicAdd(&s_cur[tile], 1u);
                s_rec[pos] = (unsigned short)((pix & (kTilePix - 1)) | ((c >> 30) << kTileShift));
                s_tile[pos] = (unsigned short)tile;
            }
        }
        __syncthreads();
        // write out: consecutive threads -> consecutive staged records -> contiguous runs inside each bucket
        const unsigned total = s_off[n_tiles - 1] + s_cnt[n_tiles - 1];
        for (unsigned i = threadIdx.x; i < total; i += blockDim.x) {
            const unsigned tile = s_tile[i];
            const unsigned slot = s_base[tile] + (i - s_off[tile]);
            const unsigned short r = s_rec[i];
            if (slot < cap) {
                buckets[(size_t)(tile * kSub + sub) * cap + slot] = r;
            } else {   // bucket full: exact fallback through the L2 reduction
                const unsigned pix = (tile << kTileShift) | (r & (kTilePix - 1));
                atomicAdd(counts + (size_t)(r >> kTileShift) * HW + pix, 1);
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024)
k_bucket_counts(const unsigned short* __restrict__ buckets, unsigned* __restrict__ cursors, unsigned cap, unsigned HW,
                int* __restrict__ counts) {
    __shared__ unsigned s_tile[2 * kTilePix];
    const unsigned tile = blockIdx.x;
    for (int i = threadIdx.x; i < 2 * kTilePix; i += blockDim.x) s_tile[i] = 0;
    __syncthreads();
    {   // two warps per sub-bucket, 4 x 16-byte loads in flight per thread
        static_assert(kSub == 16, "1024 threads = 2 warps per sub-bucket");
        const int sb = threadIdx.x >> 6, l = threadIdx.x & 63;
        const unsigned n = min(cursors[tile * kSub + sb], cap);
        const unsigned short* b = buckets + (size_t)(tile * kSub + sb) * cap;      // cap is a multiple of 8: 16-byte aligned
        const unsigned n8 = n >> 3;
        for (unsigned i0 = 0; i0 < n8; i0 += 256) {
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned i = i0 + k * 64 + l;
                v[k] = i < n8 ? ld_stream_v4(reinterpret_cast<const uint4*>(b) + i) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (i0 + k * 64 + l < n8) {
                    const unsigned w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        atomicAdd(&s_tile[w[q] & 0xfffu], 1u);        // record = pol << 11 | pixel-in-tile = index into [2][2048]
                        atomicAdd(&s_tile[(w[q] >> 16) & 0xfffu], 1u);
                    }
                }
            }
        }
        for (unsigned i = (n8 << 3) + l; i < n; i += 64) atomicAdd(&s_tile[b[i] & 0xfffu], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kSub) cursors[tile * kSub + threadIdx.x] = 0;      // leave the workspace ready for the next call
    for (int j = threadIdx.x; j < kTilePix; j += blockDim.x) {
        const unsigned pix = (tile << kTileShift) + j;
        if (pix < HW) {
            const unsigned neg = s_tile[j], pos = s_tile[kTilePix + j];
            if (neg) counts[pix] += (int)neg;
            if (pos) counts[HW + pix] += (int)pos;
        }
    }
}

static inline unsigned binned_cap(long long n, int n_tiles) {      // records per (tile, sub-bucket)
    const long long nb = (long long)n_tiles * kSub;
    long long cap = 2 * ((n + nb - 1) / nb) + 1024;
    cap = (cap + 7) & ~7ll;
    return (unsigned)cap;
}

}  // namespace evfly

using namespace evfly;

extern "C" int64_t evfly_accumulate_counts_binned_workspace_bytes(int64_t n, int H, int W) {
    if (n < 0 || H <= 0 || W <= 0) return 0;
    const int n_tiles = (int)(((long long)H * W + kTilePix - 1) / kTilePix);
    return kCursorBytes /*cursors, zero on entry*/ + (int64_t)n_tiles * kSub * binned_cap(n, n_tiles) * 2;
}

extern "C" int evfly_accumulate_counts_binned(const evfly_event* d_events, int64_t n, int H, int W, int32_t* d_counts,
                                              void* d_ws, void* stream) {
    EVFLY_REQUIRE(n >= 0 && H > 0 && W > 0 && H <= 65535 && W <= 65535 && (long long)H * W < (1ll << 30), "accumulate_counts_binned: bad n/H/W");
    EVFLY_REQUIRE(d_counts && d_ws && (d_events || n == 0), "accumulate_counts_binned: null pointer");
    if (n == 0) return EVFLY_OK;
    const unsigned HW = (unsigned)H * W;
    const int n_tiles = (int)((HW + kTilePix - 1) / kTilePix);
    EVFLY_REQUIRE(n_tiles <= kMaxTiles, "accumulate_counts_binned: frame too large (%d tiles)", n_tiles);
    const unsigned cap = binned_cap(n, n_tiles);
    unsigned* cursors = reinterpret_cast<unsigned*>(d_ws);
    unsigned short* buckets = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(d_ws) + kCursorBytes);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)n_tiles * 16 + (size_t)kChunk * 8;
    EVFLY_SMEM_ATTR(100 * 1024, k_bin_events);
    const int grid = stream_grid(n, kChunk, 6);
    k_bin_events<<<grid, 256, smem, st>>>(reinterpret_cast<const uint4*>(d_events), n, (unsigned)H, (unsigned)W, n_tiles, cap, cursors, buckets, d_counts);
    EVFLY_LAUNCHED();
    k_bucket_counts<<<n_tiles, 1024, 0, st>>>(buckets, cursors, cap, HW, d_counts);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}
