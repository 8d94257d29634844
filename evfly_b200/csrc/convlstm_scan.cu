// convlstm_scan.cu -- the ConvLSTM recurrence (1x1 kernel, no bias: convlstm.py:44-53, learner_models.py:544-546) over
// all T steps in ONE persistent cooperative kernel with the x half of the gate convolution INSIDE the step:
//
//     gates_t = x_t * W_x^T + h_{t-1} * W_h^T        (K = Cx + Ch, accumulated in TMEM)
//     c_t = sigma(f) c_{t-1} + sigma(i) tanh(g),  h_t = sigma(o) tanh(c_t)            (fp32 c, bf16 h)
//
// The round-2 timeline of the previous scan (x-gates precomputed by one big GEMM as fp32 [T*P, 4Ch], read back in the
// step epilogue) showed a step of 32 us of which 21 were the epilogue waiting for those reads chunk by chunk
// (profiles/r2_scan_timeline.txt); prefetching one chunk ahead brought 26 us. Here the x-gates never exist in memory:
// the x-part MMAs of step t+1 do not depend on h_t, so the TMA producer and the MMA issuer run them while the other
// CTAs are still finishing step t -- they fill the grid-barrier wait -- and the epilogue touches global memory only
// for c (32 B per thread and chunk, fetched one chunk ahead) and the h it publishes. The 2.7 GB fp32 x-gate tensor of a
// 1600-frame batch and the GEMM that wrote it (1.4 ms) are gone.
//
// One CTA per SM; a step's output tiles (128 rows x TN gate columns, column n = 4*channel + gate in the order i,f,o,g)
// are dealt round-robin, processed in rounds of NACC = 512 / TN tiles (one TMEM accumulator each):
//   warp 0 lane 0 : TMA producer  -- per round: x operands of its tiles; [round 0: grid barrier]; h operands
//   warp 1        : MMA issuer    -- same order; tfull[acc] is committed after a tile's h part
//   warp 2        : TMEM allocator
//   warps 4..11   : epilogue, two groups of 4 warps alternating over the round's tiles; after the step's last tile
//                   they publish h_t (fence, bar.sync, one atomic per CTA)
#include "tc_common.cuh"

namespace evfly {

struct ScanArgs {
    float* c;                  // fp32 [P, Ch] in/out
    __nv_bfloat16* h_all;      // bf16 [(T+1) * P, Ch]: block 0 = h_0, block t+1 receives h_t
    unsigned int* sync_counter;
    long long P;
    int T, Cx, Ch;
};

template <int TN>
struct ScanCfg {
    static constexpr int BM = 128, KC = 64;
    static constexpr int A_BYTES = BM * KC * 2;
    static constexpr int B_BYTES = TN * KC * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
    static constexpr int NACC = 512 / TN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(SMEM_BYTES <= 227 * 1024, "k_convlstm_scan: shared memory");
    static constexpr uint32_t SBO = 8 * KC * 2;
};

template <int TN>
__global__ void __launch_bounds__(384, 1)
k_convlstm_scan(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wx, const __grid_constant__ CUtensorMap map_h,
                const __grid_constant__ CUtensorMap map_wh, const ScanArgs p) {
    using Cfg = ScanCfg<TN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::STAGES;
    uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
    uint64_t* tempty_bar = tfull_bar + Cfg::NACC;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + Cfg::NACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (4 * p.Ch) / TN;
    const long long m_tiles = (p.P + Cfg::BM - 1) / Cfg::BM;
    const long long total_tiles = m_tiles * n_tiles;
    const int n_local = total_tiles > (long long)blockIdx.x ? (int)((total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    const int rounds = (n_local + Cfg::NACC - 1) / Cfg::NACC;
    const int kx = p.Cx / Cfg::KC, kh = p.Ch / Cfg::KC;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_wx);
        tma_prefetch_desc(&map_h);
        tma_prefetch_desc(&map_wh);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < Cfg::NACC; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        int stage = 0;
        uint32_t phase = 0;
        auto load = [&](const CUtensorMap* ma, const CUtensorMap* mb, int kc, long long a_row, int b_row) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_2d(sa, ma, &full_bar[stage], kc * Cfg::KC, (int)a_row);
            tma_load_2d(sa + Cfg::A_BYTES, mb, &full_bar[stage], kc * Cfg::KC, b_row);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        };
        for (int step = 0; step < p.T; ++step) {
            const long long row_base = (long long)step * p.P;          // x_t rows, and block `step` of h_all = h_{t-1}
            for (int r = 0; r < rounds; ++r) {
                const int j0 = r * Cfg::NACC, j1 = min(n_local, j0 + Cfg::NACC);
                for (int j = j0; j < j1; ++j) {
                    const long long tile = blockIdx.x + (long long)j * gridDim.x;
                    const long long mt = tile / n_tiles;
                    const int nt = (int)(tile - mt * n_tiles);
                    for (int kc = 0; kc < kx; ++kc) load(&map_x, &map_wx, kc, row_base + mt * Cfg::BM, nt * TN);
                }
                if (r == 0 && step > 0) {
                    // grid barrier: every CTA has published h_{step-1} (generic-proxy stores + __threadfence + atomic); acquire
                    // it, then order the async-proxy TMA reads of h after the acquire
                    const unsigned target = (unsigned)step * gridDim.x;
                    unsigned seen;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.sync_counter) : "memory");
                    } while (seen < target);
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                for (int j = j0; j < j1; ++j) {
                    const long long tile = blockIdx.x + (long long)j * gridDim.x;
                    const long long mt = tile / n_tiles;
                    const int nt = (int)(tile - mt * n_tiles);
                    for (int kc = 0; kc < kh; ++kc) load(&map_h, &map_wh, kc, row_base + mt * Cfg::BM, nt * TN);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (convergent warp, one lane issues by predicate: umma_bf16_pred) =================
        constexpr uint32_t idesc = make_idesc_bf16(Cfg::BM, TN);
        int stage = 0;
        uint32_t phase = 0;
        const uint64_t da_base = make_smem_desc(smem_u32(smem), Cfg::SBO, kLayoutSw128);
        const uint64_t db_base = make_smem_desc(smem_u32(smem) + Cfg::A_BYTES, Cfg::SBO, kLayoutSw128);
        const uint32_t elected = elect_one() ? 1u : 0u;
        auto k_blocks = [&](uint32_t tmem_d, int n_kb, bool first_overwrites) {
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t soff = (uint64_t)(stage * (Cfg::STAGE_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < Cfg::KC / 16; ++k)
                    umma_bf16_pred(tmem_d, da_base + soff + (uint64_t)(k * 2), db_base + soff + (uint64_t)(k * 2), idesc, !(first_overwrites && kb == 0 && k == 0), elected);
                umma_commit_pred(&empty_bar[stage], elected);
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        };
        uint32_t uses = 0;       // bit a = parity of the number of tiles accumulator a has held (its barriers' phase)
        for (int step = 0; step < p.T; ++step)
            for (int r = 0; r < rounds; ++r) {
                const int nj = min(n_local - r * Cfg::NACC, Cfg::NACC);
                for (int a = 0; a < nj; ++a) {
                    mbar_wait(&tempty_bar[a], ((uses >> a) & 1u) ^ 1u);      // the epilogue has drained this accumulator
                    uses ^= 1u << a;
                    tc_fence_after();
                    k_blocks(tmem_base + (uint32_t)(a * TN), kx, true);
                }
                for (int a = 0; a < nj; ++a) {
                    k_blocks(tmem_base + (uint32_t)(a * TN), kh, false);
                    umma_commit_pred(&tfull_bar[a], elected);
                }
            }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int ew = (warp - 4) & 3;           // TMEM lanes [32*ew, 32*ew+32)
        const int grp = (warp - 4) >> 2;
        uint32_t uses = 0;       // as in the MMA issuer
        for (int step = 0; step < p.T; ++step) {
            __nv_bfloat16* h_out = p.h_all + (long long)(step + 1) * p.P * p.Ch;
            for (int r = 0; r < rounds; ++r) {
                const int nj = min(n_local - r * Cfg::NACC, Cfg::NACC);
                for (int a = grp; a < nj; a += 2) {
                    const long long tile = blockIdx.x + (long long)(r * Cfg::NACC + a) * gridDim.x;
                    const long long mt = tile / n_tiles;
                    const int nt = (int)(tile - mt * n_tiles);
                    const long long m = mt * Cfg::BM + ew * 32 + lane;
                    const bool row_ok = m < p.P;
                    const long long mc = row_ok ? m : 0;                    // masked rows read row 0 and store nothing
                    float* crow = p.c + mc * (long long)p.Ch + ((nt * TN) >> 2);
                    float4 nc0 = *reinterpret_cast<const float4*>(crow), nc1 = *reinterpret_cast<const float4*>(crow + 4);   // chunk 0's c, before the wait
                    mbar_wait(&tfull_bar[a], (uses >> a) & 1u);
                    uses ^= 1u << a;
                    tc_fence_after();
#pragma unroll 1
                    for (int c0 = 0; c0 < TN; c0 += 32) {
                        const float cin[8] = {nc0.x, nc0.y, nc0.z, nc0.w, nc1.x, nc1.y, nc1.z, nc1.w};
                        if (c0 + 32 < TN) {
                            nc0 = *reinterpret_cast<const float4*>(crow + ((c0 + 32) >> 2));
                            nc1 = *reinterpret_cast<const float4*>(crow + ((c0 + 32) >> 2) + 4);
                        }
                        uint32_t v[32];
                        tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * TN + c0), v);
                        tmem_ld_wait();
                        float cn[8], hn[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float ig = fast_sigmoid(__uint_as_float(v[4 * q])), fg = fast_sigmoid(__uint_as_float(v[4 * q + 1]));
                            const float og = fast_sigmoid(__uint_as_float(v[4 * q + 2]));
                            cn[q] = fg * cin[q] + ig * fast_tanh(__uint_as_float(v[4 * q + 3]));
                            hn[q] = og * fast_tanh(cn[q]);
                        }
                        if (row_ok) {
                            float* cp = crow + (c0 >> 2);
                            *reinterpret_cast<float4*>(cp) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                            *reinterpret_cast<float4*>(cp + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                            const uint4 pk = make_uint4(cvt_bf16x2(hn[0], hn[1]), cvt_bf16x2(hn[2], hn[3]), cvt_bf16x2(hn[4], hn[5]),
                                                        cvt_bf16x2(hn[6], hn[7]));
                            *reinterpret_cast<uint4*>(h_out + m * (long long)p.Ch + ((nt * TN + c0) >> 2)) = pk;
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[a]);
                }
            }
            if (step + 1 < p.T) {
                // publish this CTA's share of h_step: stores -> gpu-scope fence -> (all 8 epilogue warps) -> one atomic
                __threadfence();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (threadIdx.x == 128) {
                    __threadfence();   // cumulative: orders the other epilogue threads' (fenced, barrier-ordered) stores too
                    atomicAdd(p.sync_counter, 1u);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int TN>
static int launch_scan(const CUtensorMap& mx, const CUtensorMap& mwx, const CUtensorMap& mh, const CUtensorMap& mwh, const ScanArgs& p, cudaStream_t st) {
    using Cfg = ScanCfg<TN>;
    EVFLY_SMEM_ATTR(Cfg::SMEM_BYTES, k_convlstm_scan<TN>);
    int per_sm = 0;
    EVFLY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_convlstm_scan<TN>, 384, Cfg::SMEM_BYTES));
    const long long resident = (long long)(per_sm > 0 ? 1 : 0) * device_sm_count();      // one CTA per SM by design (all of TMEM)
    if (resident < 1) {
        set_error("convlstm_scan_fused: the persistent kernel cannot be made resident on this device");
        return EVFLY_ERR_UNSUPPORTED;
    }
    const long long tiles = ceil_div(p.P, (long long)Cfg::BM) * ((4 * p.Ch) / TN);
    const int grid = (int)(tiles < resident ? tiles : resident);
    void* args[] = {(void*)&mx, (void*)&mwx, (void*)&mh, (void*)&mwh, (void*)&p};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_convlstm_scan<TN>, dim3(grid), dim3(384), args, Cfg::SMEM_BYTES, st);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
        cudaGetLastError();
        set_error("convlstm_scan_fused: cooperative launch of %d CTAs refused (%s)", grid, cudaGetErrorString(e));
        return EVFLY_ERR_UNSUPPORTED;
    }
    EVFLY_CUDA(e);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

}  // namespace evfly

using namespace evfly;

extern "C" int evfly_convlstm_scan_fused_bf16(const void* d_x, const void* d_wx, void* d_h_all, const void* d_wh, float* d_c, int T, int64_t P, int Cx,
                                              int Ch, void* d_sync, void* stream) {
    EVFLY_REQUIRE(d_x && d_wx && d_h_all && d_wh && d_c && d_sync && T >= 0 && P > 0, "convlstm_scan_fused_bf16: null argument");
    EVFLY_REQUIRE(Cx > 0 && Cx % 64 == 0 && Ch > 0 && Ch % 64 == 0 && 4 * Ch <= 2048, "convlstm_scan_fused_bf16: Cx %% 64 == 0, Ch %% 64 == 0, 4*Ch <= 2048 (got %d, %d)", Cx, Ch);
    EVFLY_REQUIRE((long long)(T + 1) * P < (1ll << 31), "convlstm_scan_fused_bf16: (T+1)*P must stay below 2^31 rows");
    if (T == 0) return EVFLY_OK;
    // N tile: the widest that still gives every SM a tile per step (as evfly_convlstm_scan_bf16)
    const long long m_tiles = ceil_div((long long)P, 128ll);
    int tn = 256;
    while (tn > 128 && m_tiles * ((4 * Ch) / tn) < kNumSMs) tn >>= 1;
    if ((4 * Ch) % tn) tn = 128;
    EVFLY_REQUIRE((4 * Ch) % tn == 0, "convlstm_scan_fused_bf16: 4*Ch must be a multiple of 128");
    CUtensorMap mx, mwx, mh, mwh;
    int rc = make_map_2d(&mx, d_x, (uint64_t)T * P, (uint64_t)Cx, (uint64_t)Cx, 128, 64);
    if (rc) return rc;
    rc = make_map_2d(&mwx, d_wx, (uint64_t)4 * Ch, (uint64_t)Cx, (uint64_t)Cx, (uint32_t)tn, 64);
    if (rc) return rc;
    rc = make_map_2d(&mh, d_h_all, (uint64_t)(T + 1) * P, (uint64_t)Ch, (uint64_t)Ch, 128, 64);
    if (rc) return rc;
    rc = make_map_2d(&mwh, d_wh, (uint64_t)4 * Ch, (uint64_t)Ch, (uint64_t)Ch, (uint32_t)tn, 64);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    EVFLY_CUDA(cudaMemsetAsync(d_sync, 0, 8, st));
    ScanArgs p;
    p.c = d_c;
    p.h_all = reinterpret_cast<__nv_bfloat16*>(d_h_all);
    p.sync_counter = reinterpret_cast<unsigned int*>(d_sync);
    p.P = P;
    p.T = T;
    p.Cx = Cx;
    p.Ch = Ch;
    return tn == 256 ? launch_scan<256>(mx, mwx, mh, mwh, p, st) : launch_scan<128>(mx, mwx, mh, mwh, p, st);
}
