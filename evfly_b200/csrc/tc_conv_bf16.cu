// tc_conv_bf16.cu -- L3 fast path: bf16 implicit-GEMM convolution on the 5th-gen tensor cores
// (tcgen05.mma, accumulators in TMEM, operands staged by TMA with 128B/64B swizzle).
//
// Replaces the 3x3 valid convolutions, 1x1 convolutions and 2x2-stride-2 transposed
// convolutions of OrigUNet (learner/learner_models.py:373-414, 533-583) and the ConvLSTM gate
// convolution (learner/ConvLSTM_pytorch/convlstm.py:42), i.e. ~97 % of the model's FLOPs.
//
// Layout: activations are bf16 NHWC on a PITCH-PRESERVING grid: a level's tensors all keep the
// level's input height x width as their row pitch, and a valid conv only shrinks the VALID extent
// (the last rows/cols of the grid hold don't-care values). On such a grid
//     out[m, :] = sum_{kh,kw} in[m + kh*Wp + kw, :] * W[kh,kw]          (m = flattened n,h,w)
// so a 3x3 valid conv is 9 shifted GEMMs accumulated in TMEM, and the A tile of every tap is a
// plain 2-D TMA box [128 pixels x KC channels] at row m0 + kh*Wp + kw (rows past the end are
// zero-filled by TMA). Valid outputs only ever read valid inputs, so the don't-care border never
// leaks. 2x2 max-pool compacts the valid region into the next level's grid.
//
// Kernel anatomy (one CTA = 256 threads, persistent over output tiles):
//   warp 0 lane 0 : TMA producer  (A box + B box per K-block -> smem ring, mbarrier expect_tx)
//   warp 1        : MMA issuer    (tcgen05.mma.cta_group::1.kind::f16, M=128, N=TN, K=16; one elected lane)
//   warp 2        : TMEM allocator (2 accumulator buffers of TN fp32 columns)
//   warps 4..7    : epilogue      (tcgen05.ld 32x32b -> +bias -> ReLU -> bf16 -> 16-byte stores)
#include "tc_common.cuh"
#include <string.h>

namespace evfly {


// ---------------------------------------------------------------------------------------
struct TcArgs {
    const float* bias;     // fp32 [bias_len] or nullptr; indexed by (n % bias_mod)
    const float* res;      // fp32 [M_rows, n_rows] or nullptr, added before the activation
    const __nv_bfloat16* res16;  // bf16 [M_rows, n_rows] or nullptr (token residuals)
    float* lstm_c;               // ConvLSTM fused cell update: fp32 c [M_rows, n_rows/4] in/out
    __nv_bfloat16* lstm_h;       //   and bf16 h [M_rows, n_rows/4] out; GEMM column n = 4*ch + gate (i,f,o,g)
    __nv_bfloat16* out;    // bf16 destination
    float* out_f32;        // optional fp32 destination instead (same addressing)
    long long M_rows;      // rows of the source pitch grid
    long long a_row0;      // first row of this problem inside the tensor map_a describes (ConvLSTM scan)
    // persistent scan: the kernel runs n_steps dependent problems; per step the A rows, res and lstm_h advance
    // by these strides and all CTAs meet at a grid-wide barrier (sync_counter) between steps
    int n_steps;
    long long a_row_step, res_step, lstm_h_step;
    unsigned int* sync_counter;
    int Cin, n_rows;       // K per tap; GEMM N (rows of the weight matrix)
    int taps, w_pitch;     // 1 or 9; pixels per grid row (the kh shift)
    int relu;
    long long out_ld;      // elements per destination pixel
    int out_c0;            // channel offset inside the destination pixel (concat)
    int bias_mod;
    // transposed-conv scatter: GEMM column n = phase*cout_t + co, phase = 2a+b
    int convt, Hp, Wp, valid_h, valid_w, cout_t;
    // compact output: row m = (img, ih, iw) of the source grid [N,Hp,Wp] goes to pixel (img*valid_h + ih)*valid_w + iw of a
    // [N, valid_h, valid_w] grid, rows outside valid_h x valid_w are skipped (the next conv then has no don't-care rows to compute)
    int compact;
};

template <int TN, int KC>
struct TcCfg {
    static constexpr int BM = 128;
    static constexpr int A_BYTES = BM * KC * 2;
    static constexpr int B_BYTES = TN * KC * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // as many stages as fit ~192 KB (max 10): narrow tiles have small stages and a latency-bound K loop
    static constexpr int STAGES_FIT = (192 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_FIT > 10 ? 10 : (STAGES_FIT < 3 ? 3 : STAGES_FIT);
    static constexpr int NACC = (TN <= 64) ? 4 : 2;            // TMEM accumulator buffers
    static constexpr int BIAS_FLOATS = 2048;                   // bias of every GEMM column, staged once per CTA
    static constexpr int OUT_STAGE_BYTES = 8 * 2048;           // per epilogue warp: one 32-row x 64-byte slab, transposed for contiguous stores
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + BIAS_FLOATS * 4 + OUT_STAGE_BYTES;
    static_assert(SMEM_BYTES <= 227 * 1024, "k_tc_conv_bf16: shared memory");
    static constexpr int TMEM_COLS = (NACC * TN <= 32) ? 32 : (NACC * TN <= 64 ? 64 : (NACC * TN <= 128 ? 128 : (NACC * TN <= 256 ? 256 : 512)));
    static constexpr uint32_t LAYOUT = (KC == 64) ? kLayoutSw128 : kLayoutSw64;
    static constexpr uint32_t SBO = 8 * KC * 2;  // 8 rows x row bytes
};

template <int TN, int KC>
__global__ void __launch_bounds__(384, 1)
k_tc_conv_bf16(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcArgs p) {
    using Cfg = TcCfg<TN, KC>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                       // [STAGES]
    uint64_t* empty_bar = bars + Cfg::STAGES;        // [STAGES]
    uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;                 // [NACC]
    uint64_t* tempty_bar = bars + 2 * Cfg::STAGES + Cfg::NACC;    // [NACC]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 2 * Cfg::NACC);
    float* s_bias = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);
    uint8_t* s_ostage = smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256 + Cfg::BIAS_FLOATS * 4;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m_tiles = (p.M_rows + Cfg::BM - 1) / Cfg::BM;
    const int n_tiles = (p.n_rows + TN - 1) / TN;
    const long long total_tiles = m_tiles * n_tiles;
    const int kc_per_tap = p.Cin / KC;
    const int k_blocks = p.taps * kc_per_tap;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < Cfg::NACC; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    {   // bias of GEMM column n (zero when absent / past the end): broadcast reads in the epilogue
        const int n_pad = n_tiles * TN;
        for (int i = threadIdx.x; i < n_pad; i += blockDim.x)
            s_bias[i] = (p.bias && i < p.n_rows) ? __ldg(p.bias + (i % p.bias_mod)) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        int stage = 0;
        uint32_t phase = 0;
        for (int step = 0; step < p.n_steps; ++step) {
            if (step > 0) {
                // grid barrier: every CTA's epilogue has published h_{step-1} (generic-proxy stores + __threadfence +
                // atomic); acquire it, then order the async-proxy TMA reads after the acquire
                const unsigned target = (unsigned)step * gridDim.x;
                unsigned seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.sync_counter) : "memory");
                } while (seen < target);
                asm volatile("fence.proxy.async;" ::: "memory");
            }
            const long long row_base = p.a_row0 + (long long)step * p.a_row_step;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const long long mt = tile / n_tiles;
                const int nt = (int)(tile - mt * n_tiles);
                const long long m0 = mt * Cfg::BM;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const int tap = kb / kc_per_tap, kc = kb - tap * kc_per_tap;
                    const int kh = tap / 3, kw = tap - kh * 3;  // taps == 1 -> 0,0
                    const long long row = row_base + m0 + (long long)kh * p.w_pitch + kw;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(sa, &map_a, &full_bar[stage], kc * KC, (int)row);
                    tma_load_2d(sb, &map_b, &full_bar[stage], tap * p.Cin + kc * KC, nt * TN);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // whole warp convergent, one elected lane issues; per MMA the descriptors are the stage's base descriptors plus
        // a compile-time constant in the start-address field (see tc_conv_halo.cu and profiles/r1_microbench_mma_rate.json:
        // rebuilding them from `lane == 0` costs ~75 clk per instruction, more than an N <= 128 MMA takes)
        constexpr uint32_t idesc = make_idesc_bf16(Cfg::BM, TN);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint64_t da_base = make_smem_desc(smem_u32(smem), Cfg::SBO, Cfg::LAYOUT);
        const uint64_t db_base = make_smem_desc(smem_u32(smem) + Cfg::A_BYTES, Cfg::SBO, Cfg::LAYOUT);
        for (int step = 0; step < p.n_steps; ++step)
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TN);
            for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t soff = (uint64_t)(stage * (Cfg::STAGE_BYTES >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < KC / 16; ++k)
                        umma_bf16(tmem_d, da_base + soff + (uint64_t)(k * 2), db_base + soff + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
            __syncwarp();
            if (++acc == Cfg::NACC) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        // two groups of 4 warps (4-7, 8-11) alternate over this CTA's tiles: one warp per SM sub-partition is
        // latency-bound on the epilogue of small-K problems (ViT Linear layers, ConvLSTM steps)
        const int ew = (warp - 4) & 3;  // TMEM lanes [32*ew, 32*ew+32): a warp may only touch the lanes of warp % 4
        const int grp = (warp - 4) >> 2;
        const long long n_local = total_tiles > (long long)blockIdx.x ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        for (int step = 0; step < p.n_steps; ++step) {
        const float* res_step = p.res ? p.res + (long long)step * p.res_step : nullptr;
        __nv_bfloat16* lstm_h_step = p.lstm_h ? p.lstm_h + (long long)step * p.lstm_h_step : nullptr;
        for (long long j = 0; j < n_local; ++j) {
            const long long it = (long long)step * n_local + j;      // the MMA issuer's running tile counter
            if ((int)(it & 1) != grp) continue;
            const long long tile = blockIdx.x + j * gridDim.x;
            const int acc = (int)(it % Cfg::NACC);
            const uint32_t acc_phase = (uint32_t)((it / Cfg::NACC) & 1);
            const long long mt = tile / n_tiles;
            const int nt = (int)(tile - mt * n_tiles);
            const long long m = mt * Cfg::BM + ew * 32 + lane;
            if (p.lstm_c && res_step) {
                // ===== ConvLSTM step with precomputed x-gates (the scan): the fp32 x-gates and c of the NEXT 32-column chunk are
                // fetched while the current one is computed, and chunk 0 before the accumulator is even complete -- with the
                // loads inside the chunk loop every chunk waited ~2 us for L2 and the epilogue was 21 of the 32 us of a step
                // (profiles/r2_scan_timeline.txt). Arithmetic and its order are those of the generic path below (bit-identical).
                const bool row_ok = m < p.M_rows;
                const int Ch = p.n_rows >> 2;
                const long long mc = row_ok ? m : 0;                       // masked rows read row 0 (never stored)
                const float* rrow = res_step + mc * (long long)p.n_rows + nt * TN;
                float* crow = p.lstm_c + mc * (long long)Ch + ((nt * TN) >> 2);
                float4 nres[8], nc[2];
                auto fetch = [&](int c0) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) nres[q] = *reinterpret_cast<const float4*>(rrow + c0 + 4 * q);
                    nc[0] = *reinterpret_cast<const float4*>(crow + (c0 >> 2));
                    nc[1] = *reinterpret_cast<const float4*>(crow + (c0 >> 2) + 4);
                };
                fetch(0);
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
#pragma unroll 1
                for (int c0 = 0; c0 < TN; c0 += 32) {
                    float v[32];
                    float cin[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) { v[4 * q] = nres[q].x; v[4 * q + 1] = nres[q].y; v[4 * q + 2] = nres[q].z; v[4 * q + 3] = nres[q].w; }
                    cin[0] = nc[0].x; cin[1] = nc[0].y; cin[2] = nc[0].z; cin[3] = nc[0].w; cin[4] = nc[1].x; cin[5] = nc[1].y; cin[6] = nc[1].z; cin[7] = nc[1].w;
                    if (c0 + 32 < TN) fetch(c0 + 32);
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * TN + c0), r);
                    tmem_ld_wait();
                    const float* sb = s_bias + nt * TN + c0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = (__uint_as_float(r[j]) + sb[j]) + v[j];
                    float cn[8], hn[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float ig = fast_sigmoid(v[4 * q]), fg = fast_sigmoid(v[4 * q + 1]), og = fast_sigmoid(v[4 * q + 2]);
                        cn[q] = fg * cin[q] + ig * fast_tanh(v[4 * q + 3]);
                        hn[q] = og * fast_tanh(cn[q]);
                    }
                    if (row_ok) {
                        float* cp = crow + (c0 >> 2);
                        *reinterpret_cast<float4*>(cp) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                        *reinterpret_cast<float4*>(cp + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                        uint4 pk;
                        __nv_bfloat162 t0 = __floats2bfloat162_rn(hn[0], hn[1]), t1 = __floats2bfloat162_rn(hn[2], hn[3]);
                        __nv_bfloat162 t2 = __floats2bfloat162_rn(hn[4], hn[5]), t3 = __floats2bfloat162_rn(hn[6], hn[7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                        *reinterpret_cast<uint4*>(lstm_h_step + m * (long long)Ch + ((nt * TN + c0) >> 2)) = pk;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                continue;
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            // destination pixel of this row
            bool row_ok = m < p.M_rows;
            long long dst_pix = m;
            int ih = 0, iw = 0;
            long long img = 0;
            if (p.convt || p.compact) {
                const long long hw = (long long)p.Hp * p.Wp;
                img = m / hw;
                const int rem = (int)(m - img * hw);
                ih = rem / p.Wp;
                iw = rem - ih * p.Wp;
                row_ok = row_ok && ih < p.valid_h && iw < p.valid_w;
                if (p.compact) dst_pix = (img * p.valid_h + ih) * (long long)p.valid_w + iw;
            }
#pragma unroll 1
            for (int c0 = 0; c0 < TN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * TN + c0), r);
                tmem_ld_wait();
                const int n_first = nt * TN + c0;
                if (n_first < p.n_rows) {            // warp-uniform; rows past the end / outside the valid region are masked per lane below
                    int ch0 = n_first;
                    if (p.convt) {
                        const int phs = n_first / p.cout_t;  // a 32-column slab never straddles a phase (cout_t % 32 == 0)
                        ch0 = n_first - phs * p.cout_t;
                        const int a = phs >> 1, b = phs & 1;
                        dst_pix = (img * (2 * p.valid_h) + 2 * ih + a) * (long long)(2 * p.valid_w) + 2 * iw + b;
                    }
                    float v[32];
                    const float* sb = s_bias + n_first;
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + sb[j];
                    if (res_step && row_ok) {
                        const float* rp = res_step + m * (long long)p.n_rows + n_first;
                        if (n_first + 32 <= p.n_rows && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 t = *reinterpret_cast<const float4*>(rp + 4 * q);
                                v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
                            }
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (n_first + j < p.n_rows) v[j] += rp[j];
                        }
                    }
                    if (p.res16 && row_ok) {
                        const __nv_bfloat16* rp = p.res16 + m * (long long)p.n_rows + n_first;
                        if (n_first + 32 <= p.n_rows && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 t = *reinterpret_cast<const uint4*>(rp + 8 * q);
                                const __nv_bfloat162* pt = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 f = __bfloat1622float2(pt[e]);
                                    v[8 * q + 2 * e] += f.x;
                                    v[8 * q + 2 * e + 1] += f.y;
                                }
                            }
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (n_first + j < p.n_rows) v[j] += __bfloat162float(rp[j]);
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
                    const int ncols = min(32, p.n_rows - n_first);
                    if (p.lstm_c) {
                      if (row_ok) {
                        // columns [n_first, n_first+32) = 8 channels x {i,f,o,g}: the whole cell update
                        // happens here, the gate pre-activations never go to memory (convlstm.py:44-53)
                        const int Ch = p.n_rows >> 2, ch0l = n_first >> 2;
                        float* cp = p.lstm_c + m * (long long)Ch + ch0l;
                        const float4 c_lo = *reinterpret_cast<const float4*>(cp), c_hi = *reinterpret_cast<const float4*>(cp + 4);
                        const float cin[8] = {c_lo.x, c_lo.y, c_lo.z, c_lo.w, c_hi.x, c_hi.y, c_hi.z, c_hi.w};
                        float cn[8], hn[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float ig = fast_sigmoid(v[4 * q]), fg = fast_sigmoid(v[4 * q + 1]), og = fast_sigmoid(v[4 * q + 2]);
                            cn[q] = fg * cin[q] + ig * fast_tanh(v[4 * q + 3]);
                            hn[q] = og * fast_tanh(cn[q]);
                        }
                        *reinterpret_cast<float4*>(cp) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                        *reinterpret_cast<float4*>(cp + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                        uint4 pk;
                        __nv_bfloat162 t0 = __floats2bfloat162_rn(hn[0], hn[1]), t1 = __floats2bfloat162_rn(hn[2], hn[3]);
                        __nv_bfloat162 t2 = __floats2bfloat162_rn(hn[4], hn[5]), t3 = __floats2bfloat162_rn(hn[6], hn[7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                        *reinterpret_cast<uint4*>(lstm_h_step + m * (long long)Ch + ch0l) = pk;
                      }
                    } else
                    if (p.out_f32) {
                      if (row_ok) {
                        float* o = p.out_f32 + dst_pix * p.out_ld + p.out_c0 + ch0;
                        if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                reinterpret_cast<float4*>(o)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        } else {
                            for (int j = 0; j < ncols; ++j) o[j] = v[j];
                        }
                      }
                    } else if (ncols == 32 && (p.out_ld & 7) == 0 && ((p.out_c0 + ch0) & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
                        // bf16 slab of 32 rows x 64 bytes: transposed through shared memory so that one store instruction writes
                        // 8 rows x 64 contiguous bytes (full sectors; 512 contiguous bytes when the rows are adjacent) instead of
                        // 32 scattered 16-byte pieces. Warp-uniform branch.
                        uint8_t* st = s_ostage + (warp - 4) * 2048;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 pk;
                            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[q * 8 + 0], v[q * 8 + 1]);
                            __nv_bfloat162 t1 = __floats2bfloat162_rn(v[q * 8 + 2], v[q * 8 + 3]);
                            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[q * 8 + 4], v[q * 8 + 5]);
                            __nv_bfloat162 t3 = __floats2bfloat162_rn(v[q * 8 + 6], v[q * 8 + 7]);
                            pk.x = *reinterpret_cast<uint32_t*>(&t0);
                            pk.y = *reinterpret_cast<uint32_t*>(&t1);
                            pk.z = *reinterpret_cast<uint32_t*>(&t2);
                            pk.w = *reinterpret_cast<uint32_t*>(&t3);
                            *reinterpret_cast<uint4*>(st + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = pk;
                        }
                        __syncwarp();
                        const long long row_elem = row_ok ? dst_pix * p.out_ld : -1;      // element offset of this lane's row, -1 = masked
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int rr = j * 8 + (lane >> 2), ch = lane & 3;
                            const long long re = __shfl_sync(0xffffffffu, row_elem, rr);
                            const uint4 val = *reinterpret_cast<const uint4*>(st + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
                            if (re >= 0) *reinterpret_cast<uint4*>(p.out + re + p.out_c0 + ch0 + ch * 8) = val;
                        }
                        __syncwarp();
                    } else if (row_ok) {
                        __nv_bfloat16* o = p.out + dst_pix * p.out_ld + p.out_c0 + ch0;
                        if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 pk;
                                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[q * 8 + 0], v[q * 8 + 1]);
                                __nv_bfloat162 t1 = __floats2bfloat162_rn(v[q * 8 + 2], v[q * 8 + 3]);
                                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[q * 8 + 4], v[q * 8 + 5]);
                                __nv_bfloat162 t3 = __floats2bfloat162_rn(v[q * 8 + 6], v[q * 8 + 7]);
                                pk.x = *reinterpret_cast<uint32_t*>(&t0);
                                pk.y = *reinterpret_cast<uint32_t*>(&t1);
                                pk.z = *reinterpret_cast<uint32_t*>(&t2);
                                pk.w = *reinterpret_cast<uint32_t*>(&t3);
                                reinterpret_cast<uint4*>(o)[q] = pk;
                            }
                        } else {
                            for (int j = 0; j < ncols; ++j) o[j] = __float2bfloat16_rn(v[j]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        if (p.n_steps > 1) {
            // publish this CTA's share of h_step: stores -> gpu-scope fence -> (all 8 epilogue warps) -> one atomic
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 128) {
                __threadfence();   // cumulative: orders the other epilogue threads' (fenced, barrier-ordered) stores too
                atomicAdd(p.sync_counter, 1u);
            }
            // While the other CTAs arrive and the next step's operands stream in, pull the NEXT step's fp32 gate
            // pre-activations (res = W_x * x + b, precomputed for all steps, far larger than L2) into L2: they do not
            // depend on h, and the fused cell update would otherwise wait for DRAM on the critical path of every step.
            // (Pre-issuing the weight loads before the barrier was tried as well and changed nothing.)
            if (p.res && step + 1 < p.n_steps) {
                const float* nxt = p.res + (long long)(step + 1) * p.res_step;
                const int e = threadIdx.x - 128, rr = e & 127, part = e >> 7;       // 2 threads per row
                for (long long j = 0; j < n_local; ++j) {
                    const long long tile = blockIdx.x + j * gridDim.x;
                    const long long mt = tile / n_tiles;
                    const int nt = (int)(tile - mt * n_tiles);
                    const long long m = mt * Cfg::BM + rr;
                    if (m < p.M_rows) {
                        const char* row = reinterpret_cast<const char*>(nxt + m * (long long)p.n_rows + nt * TN);
                        const int bytes = min(TN, p.n_rows - nt * TN) * 4;
                        for (int off = part * 128; off < bytes; off += 256)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(row + off));
                    }
                }
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------
PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 2-D bf16 row-major [rows, cols] (cols contiguous, row pitch ld elements), box [box_rows x box_cols]
int make_map_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                uint32_t box_cols) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return EVFLY_ERR_CUDA;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = (box_cols * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, box_cols);
        return EVFLY_ERR_CUDA;
    }
    return EVFLY_OK;
}

// cooperative = the persistent scan (n_steps > 1), whose CTAs meet at a grid-wide barrier between steps: launched with
// cudaLaunchCooperativeKernel on a grid sized from the CURRENT device (SM count x occupancy), so every CTA is
// guaranteed to be co-resident whatever the device / MIG slice / MPS limit and whatever else runs next to it
// (a plain <<<>>> launch of min(tiles, 148) CTAs spins forever when some of them cannot become resident).
// Returns EVFLY_ERR_UNSUPPORTED when the device cannot hold one CTA per launch: the caller then enqueues the steps one by one.
template <int TN, int KC>
static int launch_tc_maps(const CUtensorMap& map_a, const CUtensorMap& map_b, const TcArgs& p, cudaStream_t st, bool cooperative = false) {
    using Cfg = TcCfg<TN, KC>;
    EVFLY_SMEM_ATTR(Cfg::SMEM_BYTES, k_tc_conv_bf16<TN, KC>);
    const long long tiles = ceil_div(p.M_rows, Cfg::BM) * ceil_div(p.n_rows, TN);
    if (!cooperative) {
        const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
        k_tc_conv_bf16<TN, KC><<<grid, 384, Cfg::SMEM_BYTES, st>>>(map_a, map_b, p);
        EVFLY_LAUNCHED();
        return EVFLY_OK;
    }
    int per_sm = 0;
    EVFLY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tc_conv_bf16<TN, KC>, 384, Cfg::SMEM_BYTES));
    const long long resident = (long long)(per_sm > 0 ? 1 : 0) * device_sm_count();       // one CTA per SM by design (TMEM, 1 CTA/SM smem)
    if (resident < 1) {
        set_error("convlstm scan: the persistent kernel cannot be made resident on this device");
        return EVFLY_ERR_UNSUPPORTED;
    }
    const int grid = (int)(tiles < resident ? tiles : resident);
    void* args[] = {(void*)&map_a, (void*)&map_b, (void*)&p};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_tc_conv_bf16<TN, KC>, dim3(grid), dim3(384), args, Cfg::SMEM_BYTES, st);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
        cudaGetLastError();
        set_error("convlstm scan: cooperative launch of %d CTAs refused (%s)", grid, cudaGetErrorString(e));
        return EVFLY_ERR_UNSUPPORTED;
    }
    EVFLY_CUDA(e);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

template <int TN, int KC>
static int launch_tc(const evfly_tc_conv_args& a, const TcArgs& p, cudaStream_t st) {
    using Cfg = TcCfg<TN, KC>;
    CUtensorMap map_a, map_b;
    int rc = make_map_2d(&map_a, a.x, (uint64_t)a.M_rows, (uint64_t)a.Cin, (uint64_t)a.Cin, Cfg::BM, KC);
    if (rc) return rc;
    rc = make_map_2d(&map_b, a.w, (uint64_t)a.n_rows, (uint64_t)a.taps * a.Cin, (uint64_t)a.taps * a.Cin, TN, KC);
    if (rc) return rc;
    return launch_tc_maps<TN, KC>(map_a, map_b, p, st);
}

}  // namespace evfly

using namespace evfly;

extern "C" int evfly_tc_conv_bf16(const evfly_tc_conv_args* args, void* stream) {
    EVFLY_REQUIRE(args, "tc_conv_bf16: null args");
    const evfly_tc_conv_args a = *args;
    EVFLY_REQUIRE(a.x && a.w && (a.out || a.out_f32 || a.lstm_c), "tc_conv_bf16: null tensor");
    EVFLY_REQUIRE((a.lstm_c == nullptr) == (a.lstm_h == nullptr), "tc_conv_bf16: lstm_c / lstm_h go together");
    EVFLY_REQUIRE(!a.lstm_c || (a.n_rows % 32 == 0 && !a.convt && !a.relu), "tc_conv_bf16: fused ConvLSTM needs n_rows %% 32 == 0");
    EVFLY_REQUIRE(a.M_rows > 0 && a.M_rows < (1ll << 31) && a.Cin > 0 && a.n_rows > 0 && a.n_rows <= 2048, "tc_conv_bf16: bad shape (n_rows <= 2048)");
    EVFLY_REQUIRE(a.taps == 1 || a.taps == 9, "tc_conv_bf16: taps must be 1 or 9");
    EVFLY_REQUIRE(a.Cin % 32 == 0, "tc_conv_bf16: Cin must be a multiple of 32 (got %d)", a.Cin);
    EVFLY_REQUIRE((reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0, "tc_conv_bf16: operands must be 16-byte aligned");
    EVFLY_REQUIRE(!(a.convt && (a.res_f32 || a.res_bf16)), "tc_conv_bf16: residuals are not supported with convt");
    EVFLY_REQUIRE(!(a.flags & EVFLY_TC_COMPACT) || (!a.convt && !a.lstm_c && !a.res_f32 && !a.res_bf16 && a.Hp > 0 && a.Wp > 0 && a.valid_h >= 1 && a.valid_h <= a.Hp &&
                                                    a.valid_w >= 1 && a.valid_w <= a.Wp && a.M_rows % ((int64_t)a.Hp * a.Wp) == 0),
                  "tc_conv_bf16: EVFLY_TC_COMPACT needs Hp, Wp, valid_h <= Hp, valid_w <= Wp, M_rows = N*Hp*Wp and no residual / convt / lstm");
    EVFLY_REQUIRE(!a.convt || (a.taps == 1 && a.cout_t % 32 == 0 && a.n_rows == 4 * a.cout_t && a.Hp > 0 && a.Wp > 0 && a.valid_h <= a.Hp && a.valid_w <= a.Wp),
                  "tc_conv_bf16: bad transposed-conv arguments");
    TcArgs p;
    p.bias = a.bias;
    p.res = a.res_f32;
    p.res16 = reinterpret_cast<const __nv_bfloat16*>(a.res_bf16);
    p.lstm_c = a.lstm_c;
    p.lstm_h = reinterpret_cast<__nv_bfloat16*>(a.lstm_h);
    p.out = reinterpret_cast<__nv_bfloat16*>(a.out);
    p.out_f32 = a.out_f32;
    p.M_rows = a.M_rows;
    p.a_row0 = 0;
    p.n_steps = 1;
    p.a_row_step = p.res_step = p.lstm_h_step = 0;
    p.sync_counter = nullptr;
    p.Cin = a.Cin;
    p.n_rows = a.n_rows;
    p.taps = a.taps;
    p.w_pitch = a.w_pitch;
    p.relu = a.relu;
    p.out_ld = a.out_ld;
    p.out_c0 = a.out_c0;
    p.bias_mod = a.convt ? a.cout_t : a.n_rows;
    p.convt = a.convt;
    p.Hp = a.Hp;
    p.Wp = a.Wp;
    p.valid_h = a.valid_h;
    p.valid_w = a.valid_w;
    p.cout_t = a.cout_t;
    p.compact = (int)(a.flags & EVFLY_TC_COMPACT);
    cudaStream_t st = (cudaStream_t)stream;
    const bool kc64 = (a.Cin % 64 == 0);
    // N tile: the whole weight matrix when it fits 256 accumulator columns, else 256-wide tiles;
    // small-M problems (ConvLSTM step, deep UNet levels at small batch) take narrower tiles so that
    // m_tiles * n_tiles covers the 148 SMs.
    const int n = a.n_rows;
    const long long m_tiles = ceil_div(a.M_rows, 128);
    int tn = n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256));
    while (tn > 32 && m_tiles * ceil_div(n, tn) < kNumSMs) tn >>= 1;
    if (kc64) {
        if (tn == 32) return launch_tc<32, 64>(a, p, st);
        if (tn == 64) return launch_tc<64, 64>(a, p, st);
        if (tn == 128) return launch_tc<128, 64>(a, p, st);
        return launch_tc<256, 64>(a, p, st);
    } else {
        if (tn == 32) return launch_tc<32, 32>(a, p, st);
        if (tn == 64) return launch_tc<64, 32>(a, p, st);
        if (tn == 128) return launch_tc<128, 32>(a, p, st);
        return launch_tc<256, 32>(a, p, st);
    }
}

// ConvLSTM (1x1 kernel) scan over T steps in ONE call: the tensor maps are encoded once and the T fused
// step kernels (h-gates GEMM + x-gates + cell update, see lstm_c above) are enqueued back to back from C++.
// d_h_all bf16 [(T+1)*P, Ch]: block 0 holds h_0 on entry, block t+1 receives h_t.
extern "C" int evfly_convlstm_scan_bf16(void* d_h_all, const void* d_wh, const float* d_gx, float* d_c, int T, int64_t P, int Ch,
                                        void* d_sync, void* stream) {
    EVFLY_REQUIRE(d_h_all && d_wh && d_gx && d_c && T >= 0 && P > 0 && Ch > 0 && Ch % 64 == 0 && 4 * Ch <= 2048, "convlstm_scan_bf16: bad argument (Ch %% 64 == 0, 4*Ch <= 2048)");
    if (T == 0) return EVFLY_OK;
    constexpr int KC = 64;
    // N tile as in evfly_tc_conv_bf16: the widest that still gives >= 148 tiles per step
    const long long m_tiles = ceil_div(P, 128);
    int tn = 256;
    while (tn > 32 && m_tiles * ceil_div(4 * Ch, tn) < kNumSMs) tn >>= 1;
    CUtensorMap map_a, map_b;
    int rc = make_map_2d(&map_a, d_h_all, (uint64_t)(T + 1) * P, (uint64_t)Ch, (uint64_t)Ch, 128, KC);
    if (rc) return rc;
    rc = make_map_2d(&map_b, d_wh, (uint64_t)4 * Ch, (uint64_t)Ch, (uint64_t)Ch, (uint32_t)tn, KC);
    if (rc) return rc;
    TcArgs p;
    memset(&p, 0, sizeof(p));
    p.M_rows = P;
    p.Cin = Ch;
    p.n_rows = 4 * Ch;
    p.taps = 1;
    p.bias_mod = 4 * Ch;
    p.out_ld = 4 * Ch;
    p.lstm_c = d_c;
    __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(d_h_all);
    cudaStream_t st = (cudaStream_t)stream;
    p.n_steps = 1;
    if (d_sync && T > 1) {
        // persistent: ONE cooperative launch walks all T steps; CTAs meet at a grid-wide barrier between steps
        EVFLY_CUDA(cudaMemsetAsync(d_sync, 0, 8, st));
        TcArgs q = p;
        q.n_steps = T;
        q.a_row0 = 0;
        q.a_row_step = P;
        q.res = d_gx;
        q.res_step = (long long)P * 4 * Ch;
        q.lstm_h = h + (long long)P * Ch;
        q.lstm_h_step = (long long)P * Ch;
        q.sync_counter = reinterpret_cast<unsigned int*>(d_sync);
        rc = tn == 32 ? launch_tc_maps<32, KC>(map_a, map_b, q, st, true)
           : tn == 64 ? launch_tc_maps<64, KC>(map_a, map_b, q, st, true)
           : tn == 128 ? launch_tc_maps<128, KC>(map_a, map_b, q, st, true)
                       : launch_tc_maps<256, KC>(map_a, map_b, q, st, true);
        if (rc != EVFLY_ERR_UNSUPPORTED) return rc;
        // co-residency cannot be guaranteed here: fall through to one launch per step (same arithmetic, bit-identical)
    }
    for (int t = 0; t < T; ++t) {
        p.a_row0 = (long long)t * P;
        p.res = d_gx + (long long)t * P * 4 * Ch;
        p.lstm_h = h + (long long)(t + 1) * P * Ch;
        rc = tn == 32 ? launch_tc_maps<32, KC>(map_a, map_b, p, st)
           : tn == 64 ? launch_tc_maps<64, KC>(map_a, map_b, p, st)
           : tn == 128 ? launch_tc_maps<128, KC>(map_a, map_b, p, st)
                       : launch_tc_maps<256, KC>(map_a, map_b, p, st);
        if (rc) return rc;
    }
    return EVFLY_OK;
}
