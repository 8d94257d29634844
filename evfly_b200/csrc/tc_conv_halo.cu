// tc_conv_halo.cu -- 3x3 valid convolution for the SMALL-CHANNEL UNet layers (Cin, Cout <= 64:
// unet_e12, e21, e22, d32, d41, d42 -- the big-spatial half of the FLOPs) on the tensor cores with the
// input halo reused from shared memory.
//
// k_tc_conv_bf16 feeds every tap of every tile with its own TMA box, 9x the input traffic; at 32-64
// channels a tap's MMA work (32-128 cycles) is far below the cost of issuing its 128 TMA row
// requests, so those layers were TMA-request bound (168-400 TFLOP/s). Here one output tile is a
// 16x8 pixel block; its 18x10 halo is ONE 4-D TMA box [1, 18, 10, Cin] (180 rows of 64/128 B in
// shared memory, pixel x fastest) and each of the 9 taps is read by tcgen05.mma through a descriptor
// that STARTS at row (kh*10 + kw) with an 8-row-group stride of 10 rows (SBO = 10 * row bytes): the
// 128 A rows of tap (kh,kw) are exactly the pixels (r+kh, c+kw), r<16, c<8. The swizzle is applied on
// absolute smem address bits, so shifted starts are exact (evfly_tc_shift_probe). The weights
// (9 x Cout x Cin bf16, <= 72 KB) are loaded once per CTA and stay resident.
//
// warp 0 lane 0: TMA producer (halo ring) | warp 1 lane 0: MMA issuer | warp 2: TMEM alloc |
// warps 4-11: epilogue in two groups of 4 warps that alternate tiles (+bias, ReLU, bf16, masked 64/128-byte
// stores to the valid pixels): one warp per SM sub-partition was latency-bound at ~1 us per tile
//
// Round 2 (profiles/r2_exp_halo_roles.txt: with each role in turn reduced to its barrier traffic, the 32-channel layers
// were bound by the INSTRUCTION COUNT of the epilogue and of the tile bookkeeping, ~3500 warp instructions per 128-pixel
// tile against 720 clk of MMA): the epilogue converts with cvt.rn[.relu].bf16x2 (one instruction per channel pair, ReLU
// included), pools from the staging rows it writes anyway (4 LDS + 12 packed max instead of 64 shuffles + 64 max), no role
// divides per tile (TileWalk), and the layers whose weights + halos fit twice run 2 CTAs per SM so that one CTA's
// handshake latency is covered by the other's MMAs.
#include "tc_common.cuh"

namespace evfly {

struct HaloArgs {
    const float* bias;
    __nv_bfloat16* out;   // [N, Hp, Wp, COUT], same pitch as the input
    __nv_bfloat16* pool_out;   // optional fused MaxPool2d(2): [N, Hp2, Wp2, COUT], valid (out_vh/2) x (out_vw/2)
    int Hp2, Wp2;
    int pad;              // 1: padding=1 conv on a dense tensor (Hp x Wp all valid): the halo box starts at (-1,-1) and TMA zero-fills outside the image
    int N, Hp, Wp, out_vh, out_vw;
    int oHp, oWp;         // pitch of `out` / `out1`: the input's (Hp, Wp), or out_vh x out_vw for a compact output grid
    int tiles_x, tiles_y;
    uint32_t magic_x;     // floor(2^32 / tiles_x) + 1 when tiles_x * tiles_y < 65536 (then umulhi(rem, magic_x) == rem / tiles_x), else 0
    int relu;
    // optional fused 1x1 output conv to ONE channel (unet_out after unet_d42, learner_models.py:583): out1[n, oh, ow] =
    // b1 + sum_c bf16(act(conv)[c]) * w1[c] in fp32 on the same pitch grid; with it `out` may be null (the COUT-channel
    // activation is then never written)
    const float* w1;      // [COUT] (bf16-rounded values as fp32), or nullptr
    const float* b1;      // [1]
    float* out1;          // [N, Hp, Wp] fp32
    // optional: `out` will only be sampled by evfly_resize_bilinear_nhwc_bf16 to a height of skip_OH (the decoder's interp skip,
    // learner_models.py:512-519): bit r set = output row r is read by that resize; rows with a clear bit are not written
    uint32_t row_mask[16];
    int use_rows;
};

__device__ __forceinline__ bool row_wanted(const HaloArgs& p, int r) { return !p.use_rows || ((p.row_mask[r >> 5] >> (r & 31)) & 1u); }

template <int CIN, int COUT>
struct HaloCfg {
    static constexpr int ROW_B = CIN * 2;                       // bytes per pixel row in smem
    static constexpr int HALO_ROWS = 18 * 10;
    static constexpr int HALO_BYTES = ((HALO_ROWS * ROW_B + 1023) / 1024) * 1024;
    static constexpr int W_TAP_BYTES = COUT * ROW_B;
    static constexpr int W_BYTES = 9 * W_TAP_BYTES;
    // 2 CTAs per SM where weights + halo ring + staging fit in half the shared memory (32->32, 32->64, 64->32): a second
    // CTA's MMAs fill the tensor pipe while the first waits on its own producer/epilogue handshakes
    static constexpr int MINB = (CIN == 32 || (CIN == 64 && COUT == 32)) ? 2 : 1;
    static constexpr int STAGES = (MINB == 2) ? (CIN == 32 ? 3 : 2) : (COUT > 64) ? 3 : 4;   // 64->128: 144 KB of resident weights leave room for 3 halos
    static constexpr int NACC = 4;
    static constexpr int TMEM_COLS = (NACC * COUT <= 128) ? 128 : (NACC * COUT <= 256) ? 256 : 512;
    // epilogue staging (8 warps x 32 pixels x COUT bf16) so that global stores are 512-byte contiguous; not for COUT = 128
    static constexpr int OUT_STAGE_BYTES = (COUT <= 64) ? 8 * 32 * COUT * 2 : 0;
    static constexpr int SMEM_BYTES = W_BYTES + STAGES * HALO_BYTES + 1024 /*alignment slack*/ + 2048 /*barriers, bias*/ + OUT_STAGE_BYTES;
    static_assert(SMEM_BYTES <= 227 * 1024, "halo conv: weights + halo stages exceed shared memory");
    static_assert(MINB == 1 || 2 * (SMEM_BYTES + 1024) <= 227 * 1024, "halo conv: two CTAs per SM do not fit");
    static constexpr uint32_t LAYOUT = (CIN == 64) ? kLayoutSw128 : kLayoutSw64;
    static constexpr uint32_t SBO_A = 10 * ROW_B;               // next 8-pixel group = next output row = 10 halo rows
    static constexpr uint32_t SBO_B = 8 * ROW_B;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// explicit shared-space 16-byte accesses: through generic pointers the table copy compiled to LD.E / ST.E, whose stores
// were the producers' top stall (ncu source page, profiles/r2_ncu_stem_e12.txt)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Position of a role in its strided walk over the tiles (tile = n * tiles_per_img + rem) without a division per tile:
// the stride is split once into whole images and a remainder.
struct TileWalk {
    int n, rem, step_n, step_rem, tpi;
    __device__ __forceinline__ TileWalk(long long first, long long step, int tiles_per_img) {
        tpi = tiles_per_img;
        n = (int)(first / tiles_per_img);
        rem = (int)(first - (long long)n * tiles_per_img);
        step_n = (int)(step / tiles_per_img);
        step_rem = (int)(step - (long long)step_n * tiles_per_img);
    }
    __device__ __forceinline__ void next() {
        n += step_n;
        rem += step_rem;
        if (rem >= tpi) { rem -= tpi; ++n; }
    }
    __device__ __forceinline__ void yx(const HaloArgs& p, int& ty, int& tx) const {
        ty = p.magic_x ? (int)__umulhi((unsigned)rem, p.magic_x) : rem / p.tiles_x;
        tx = rem - ty * p.tiles_x;
    }
};

// Epilogue shared by the two halo kernels: warps 4-7 take the even local tiles, warps 8-11 the odd ones.
template <int COUT, int NACC, bool STAGE_OUT>
__device__ __forceinline__ void halo_epilogue(const HaloArgs& p, int warp, int lane, uint32_t tmem_base, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                              const float* s_bias, uint8_t* s_ostage, int total_tiles, int tiles_per_img) {
    // ================= epilogue: group 0 = warps 4-7 (even local tiles), group 1 = warps 8-11 (odd) =====
    const int ew = (warp - 4) & 3;           // TMEM lanes [32*ew, 32*ew+32) (a warp may only touch lanes of warp%4)
    const int grp = (warp - 4) >> 2;
    const int row = ew * 32 + lane;          // tile pixel: r = row / 8, c = row % 8
    const int r = row >> 3, c = row & 7;
    constexpr bool kStage = STAGE_OUT;
    constexpr int kCP = COUT / 8;            // 16-byte chunks per pixel
    const uint32_t my_stage = smem_u32(s_ostage) + (warp - 4) * (32 * COUT * 2);      // shared-space address: STS / LDS, not generic ST / LD
    const uint32_t my_row = my_stage + lane * (COUT * 2);
    const uint32_t my_swz = kCP == 4 ? ((lane >> 1) & 3) : (lane & 7);
    const float4* s_bias4 = reinterpret_cast<const float4*>(s_bias);                   // broadcast reads: registers are what two CTAs per SM are short of
    TileWalk tw((long long)blockIdx.x + (long long)grp * gridDim.x, 2ll * gridDim.x, tiles_per_img);
    int it = grp;
    for (long long tile = (long long)blockIdx.x + (long long)grp * gridDim.x; tile < total_tiles; tile += 2ll * gridDim.x, it += 2, tw.next()) {
        const int acc = it & (NACC - 1);
        const uint32_t acc_phase = (uint32_t)(it / NACC) & 1u;
        const int n = tw.n;
        int ty, tx;
        tw.yx(p, ty, tx);
        const int oh = ty * 16 + r, ow = tx * 8 + c;
        const bool ok = oh < p.out_vh && ow < p.out_vw;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        float dot1 = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * COUT + c0), v);
            tmem_ld_wait();
            uint32_t pk[16];                 // channels c0 + 2i, c0 + 2i + 1
            if (p.relu) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b = s_bias4[(c0 >> 2) + i];
                    pk[2 * i] = cvt_relu_bf16x2(__uint_as_float(v[4 * i]) + b.x, __uint_as_float(v[4 * i + 1]) + b.y);
                    pk[2 * i + 1] = cvt_relu_bf16x2(__uint_as_float(v[4 * i + 2]) + b.z, __uint_as_float(v[4 * i + 3]) + b.w);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b = s_bias4[(c0 >> 2) + i];
                    pk[2 * i] = cvt_bf16x2(__uint_as_float(v[4 * i]) + b.x, __uint_as_float(v[4 * i + 1]) + b.y);
                    pk[2 * i + 1] = cvt_bf16x2(__uint_as_float(v[4 * i + 2]) + b.z, __uint_as_float(v[4 * i + 3]) + b.w);
                }
            }
            if (p.w1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    dot1 = fmaf(__uint_as_float(pk[i] << 16), __ldg(p.w1 + c0 + 2 * i), dot1);
                    dot1 = fmaf(__uint_as_float(pk[i] & 0xffff0000u), __ldg(p.w1 + c0 + 2 * i + 1), dot1);
                }
            }
            if (!p.out) continue;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 val = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                if constexpr (kStage) {
                    // this pixel's chunk (c0/8 + q) -> staging row `lane`, chunk position XOR-swizzled (conflict-free)
                    sts128(my_row + (((uint32_t)((c0 >> 3) + q) ^ my_swz) << 4), val);
                } else {
                    if (ok && row_wanted(p, oh)) reinterpret_cast<uint4*>(p.out + (((long long)n * p.oHp + oh) * p.oWp + ow) * COUT + c0)[q] = val;
                    if (p.pool_out) {
                        // fused 2x2 max-pool: the partners of pixel (r, c) are lanes ^1 (column) and ^8 (row) of the same warp; the
                        // lane with even r and even c writes. max commutes with the (monotonic) bf16 rounding, so the result
                        // is bit-identical to pooling the stored tensor.
                        uint4 m = max4_bf16x2(val, make_uint4(__shfl_xor_sync(0xffffffffu, val.x, 1), __shfl_xor_sync(0xffffffffu, val.y, 1),
                                                              __shfl_xor_sync(0xffffffffu, val.z, 1), __shfl_xor_sync(0xffffffffu, val.w, 1)));
                        m = max4_bf16x2(m, make_uint4(__shfl_xor_sync(0xffffffffu, m.x, 8), __shfl_xor_sync(0xffffffffu, m.y, 8),
                                                      __shfl_xor_sync(0xffffffffu, m.z, 8), __shfl_xor_sync(0xffffffffu, m.w, 8)));
                        if (((lane & 9) == 0) && oh + 1 < p.out_vh && ow + 1 < p.out_vw)
                            reinterpret_cast<uint4*>(p.pool_out + (((long long)n * p.Hp2 + (oh >> 1)) * p.Wp2 + (ow >> 1)) * COUT + c0)[q] = m;
                    }
                }
            }
        }
        if (p.w1 && ok) p.out1[((long long)n * p.oHp + oh) * p.oWp + ow] = dot1 + __ldg(p.b1);
        if constexpr (kStage) if (p.out) {
            __syncwarp();
            if (p.pool_out) {
                // fused 2x2 max-pool from the staging rows: the warp's 4 x 8 pixels hold 2 x 4 pooled pixels of kCP chunks; one
                // (pooled pixel, chunk) per lane and pass. Lanes of odd pooled pixels read the odd column first, so that the eight
                // lanes of a shared-memory phase cover all eight 16-byte bank groups (COUT = 32: two pixels per 128 bytes). max
                // commutes with the (monotonic) bf16 rounding: bit-identical to pooling the stored tensor.
#pragma unroll
                for (int j = 0; j < kCP / 4; ++j) {
                    const int id = j * 32 + lane, pp = id / kCP, ch = id % kCP;
                    const int pr = pp >> 2, pc = pp & 3, f = pp & 1;
                    const int px0 = pr * 16 + pc * 2;
                    auto at = [&](int px) { return lds128(my_stage + px * (COUT * 2) + (((uint32_t)ch ^ (uint32_t)(kCP == 4 ? ((px >> 1) & 3) : (px & 7))) << 4)); };
                    const uint4 a0 = at(px0 + f), a1 = at(px0 + (f ^ 1)), a2 = at(px0 + 8 + f), a3 = at(px0 + 8 + (f ^ 1));
                    const uint4 m = max4_bf16x2(max4_bf16x2(a0, a1), max4_bf16x2(a2, a3));
                    const int ph = ty * 8 + ew * 2 + pr, pw = tx * 4 + pc;
                    if (2 * ph + 1 < p.out_vh && 2 * pw + 1 < p.out_vw)
                        *reinterpret_cast<uint4*>(p.pool_out + (((long long)n * p.Hp2 + ph) * p.Wp2 + pw) * COUT + ch * 8) = m;
                }
            }
            // write-out: instruction j stores chunks [32j, 32j+32) of the warp's 32 pixels = 512 contiguous bytes
            // (8 pixels of one output row are adjacent in the NHWC grid)
#pragma unroll
            for (int j = 0; j < kCP; ++j) {
                const int id = j * 32 + lane, px = id / kCP, ch = id % kCP;
                const int poh = ty * 16 + ew * 4 + (px >> 3), pow_ = tx * 8 + (px & 7);
                if (poh < p.out_vh && pow_ < p.out_vw && row_wanted(p, poh)) {
                    const uint4 val = lds128(my_stage + px * (COUT * 2) + ((ch ^ (kCP == 4 ? ((px >> 1) & 3) : (px & 7))) << 4));
                    *reinterpret_cast<uint4*>(p.out + (((long long)n * p.oHp + poh) * p.oWp + pow_) * COUT + ch * 8) = val;
                }
            }
            __syncwarp();      // the staging rows are rewritten by the next tile of this warp
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(384, (HaloCfg<CIN, COUT>::MINB))
k_tc_conv3x3_halo(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const HaloArgs p) {
    using Cfg = HaloCfg<CIN, COUT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;                                  // [9][COUT][CIN] swizzled per tap
    uint8_t* s_halo = smem + Cfg::W_BYTES;                // [STAGES][180][CIN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_halo + Cfg::STAGES * Cfg::HALO_BYTES);
    uint64_t* full_bar = bars;                            // [STAGES]
    uint64_t* empty_bar = bars + Cfg::STAGES;             // [STAGES]
    uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;         // [NACC]
    uint64_t* tempty_bar = tfull_bar + Cfg::NACC;         // [NACC]
    uint64_t* w_bar = tempty_bar + Cfg::NACC;             // [1]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_bar + 1);
    float* s_bias = reinterpret_cast<float*>(bars + 32);  // 16-byte aligned: the epilogue reads it as float4  // [COUT]
    uint8_t* s_ostage = s_halo + Cfg::STAGES * Cfg::HALO_BYTES + 2048;   // [8 warps][32 px][COUT] bf16 (16-byte chunks XOR-swizzled)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int total_tiles = tiles_per_img * p.N;      // host guarantees < 2^31

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
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
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    if (threadIdx.x < COUT) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        mbar_expect_tx(w_bar, Cfg::W_BYTES);
        for (int t = 0; t < 9; ++t) tma_load_2d(s_w + t * Cfg::W_TAP_BYTES, &map_w, w_bar, t * CIN, 0);
        int stage = 0;
        uint32_t phase = 0;
        TileWalk tw(blockIdx.x, gridDim.x, tiles_per_img);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, tw.next()) {
            int ty, tx;
            tw.yx(p, ty, tx);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], Cfg::HALO_ROWS * Cfg::ROW_B);
            tma_load_4d(s_halo + stage * Cfg::HALO_BYTES, &map_x, &full_bar[stage], 0, tx * 8 - p.pad, ty * 16 - p.pad, tw.n);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp walks the tile loop convergently and ONE elected lane issues, by predicate (umma_bf16_pred): the
        // descriptors of a tile differ from the stage's base descriptor only in the start-address field (bits [0,14), units
        // of 16 B) by compile-time constants and stay in uniform registers. (Issuing from `lane == 0` with the descriptors
        // rebuilt per MMA took ~75 clk per instruction, from an `if (elect_one())` region ~26 clk of R2UR moves per MMA; the
        // tensor pipe needs ~40 at N <= 32, profiles/r1_microbench_mma_rate.json.)
        constexpr uint32_t idesc = make_idesc_bf16(128, COUT);
        mbar_wait(w_bar, 0);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        const uint64_t dw0 = make_smem_desc(smem_u32(s_w), Cfg::SBO_B, Cfg::LAYOUT);
        const uint64_t dh0 = make_smem_desc(smem_u32(s_halo), Cfg::SBO_A, Cfg::LAYOUT);
        const uint32_t elected = elect_one() ? 1u : 0u;      // one lane issues every MMA and commit of this CTA (umma_bf16_pred)
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * COUT);
            const uint64_t da0 = dh0 + (uint64_t)(stage * (Cfg::HALO_BYTES >> 4));
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                for (int k = 0; k < CIN / 16; ++k) {
                    const uint64_t da = da0 + (uint64_t)((((tap / 3) * 10 + (tap % 3)) * Cfg::ROW_B + k * 32) >> 4);
                    const uint64_t db = dw0 + (uint64_t)((tap * Cfg::W_TAP_BYTES + k * 32) >> 4);
                    umma_bf16_pred(tmem_d, da, db, idesc, (tap | k) != 0, elected);
                }
            }
            umma_commit_pred(&empty_bar[stage], elected);
            umma_commit_pred(&tfull_bar[acc], elected);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            if (++acc == Cfg::NACC) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        halo_epilogue<COUT, Cfg::NACC, (Cfg::OUT_STAGE_BYTES > 0)>(p, warp, lane, tmem_base, tfull_bar, tempty_bar, s_bias, s_ostage, total_tiles, tiles_per_img);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------
// Cin = 128 layers (e32, d32, d41, e41): the weights (9 x COUT x 128 bf16, up to 590 KB) cannot stay in shared memory,
// but the input halo still can. Per 16x8 output tile the halo is loaded ONCE as two 64-channel chunks (two 4-D TMA
// boxes of 180 rows x 128 B, double-buffered across tiles) and the 18 weight blocks [COUT x 64] of (chunk, tap) stream
// through a small ring. Against the tap-streaming kernel (A tile + B block per K-block) the L2 -> SM traffic per tile
// drops from 18 x (16 + B) KB to 46 + 18 x B KB, which is what bounds these layers (they sat at 560-880 TFLOP/s with
// the tensor pipe a third busy).
// ---------------------------------------------------------------------------------------
template <int COUT>
struct HaloWsCfg {
    static constexpr int CHUNK_BYTES = ((180 * 128 + 1023) / 1024) * 1024;      // one 64-channel halo chunk, 128-byte rows
    static constexpr int A_BYTES = 2 /*tile buffers*/ * 2 /*chunks*/ * CHUNK_BYTES;
    static constexpr int B_BLOCK = COUT * 128;                                   // [COUT rows][64 ch] bf16
    static constexpr int NB = (COUT <= 128) ? 6 : 3;                             // weight ring slots
    static constexpr int NACC = (COUT <= 128) ? 4 : 2;
    static constexpr int TMEM_COLS = 512;
    static constexpr int OUT_STAGE_BYTES = (COUT <= 64) ? 8 * 32 * COUT * 2 : 0;
    static constexpr int SMEM_BYTES = A_BYTES + NB * B_BLOCK + 1024 + 2048 + OUT_STAGE_BYTES;
    static_assert(SMEM_BYTES <= 227 * 1024, "weight-streaming halo conv: shared memory");
    static constexpr uint32_t SBO_A = 10 * 128, SBO_B = 8 * 128;
};

template <int COUT>
__global__ void __launch_bounds__(384, 1)
k_tc_conv3x3_halo_ws(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const HaloArgs p) {
    using Cfg = HaloWsCfg<COUT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_a = smem;                                  // [2 buffers][2 chunks][180 rows][128 B]
    uint8_t* s_b = smem + Cfg::A_BYTES;                   // [NB][COUT][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_b + Cfg::NB * Cfg::B_BLOCK);
    uint64_t* a_full = bars;                              // [2]
    uint64_t* a_empty = bars + 2;                         // [2]
    uint64_t* b_full = bars + 4;                          // [NB]
    uint64_t* b_empty = b_full + Cfg::NB;                 // [NB]
    uint64_t* tfull_bar = b_empty + Cfg::NB;              // [NACC]
    uint64_t* tempty_bar = tfull_bar + Cfg::NACC;         // [NACC]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + Cfg::NACC);
    float* s_bias = reinterpret_cast<float*>(bars + 64);               // [COUT], 16-byte aligned: the epilogue reads it as float4
    uint8_t* s_ostage = s_b + Cfg::NB * Cfg::B_BLOCK + 2048;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int total_tiles = tiles_per_img * p.N;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < Cfg::NB; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int a = 0; a < Cfg::NACC; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        int slot = 0;
        uint32_t b_phase = 0;
        int it = 0;
        TileWalk tw(blockIdx.x, gridDim.x, tiles_per_img);      // load_halo is called once per tile, in tile order
        auto load_halo = [&](int tile, int j) {       // j = running local tile index -> buffer j & 1
            const int buf = j & 1;
            const int n = tw.n;
            int ty, tx;
            tw.yx(p, ty, tx);
            tw.next();
            mbar_wait(&a_empty[buf], (((uint32_t)j >> 1) & 1u) ^ 1u);
            mbar_expect_tx(&a_full[buf], 2 * 180 * 128);
            tma_load_4d(s_a + (buf * 2 + 0) * Cfg::CHUNK_BYTES, &map_x, &a_full[buf], 0, tx * 8 - p.pad, ty * 16 - p.pad, n);
            tma_load_4d(s_a + (buf * 2 + 1) * Cfg::CHUNK_BYTES, &map_x, &a_full[buf], 64, tx * 8 - p.pad, ty * 16 - p.pad, n);
        };
        if ((int)blockIdx.x < total_tiles) load_halo(blockIdx.x, 0);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            for (int blk = 0; blk < 18; ++blk) {                                                // (chunk, tap) = (blk / 9, blk % 9)
                // prefetch the next tile's halo once this tile's first weight blocks are on their way (its buffer is
                // released by the previous tile's last MMA, which has retired by then)
                if (blk == 9 && tile + (int)gridDim.x < total_tiles) load_halo(tile + gridDim.x, it + 1);
                mbar_wait(&b_empty[slot], b_phase ^ 1);
                mbar_expect_tx(&b_full[slot], Cfg::B_BLOCK);
                tma_load_2d(s_b + slot * Cfg::B_BLOCK, &map_w, &b_full[slot], (blk % 9) * 128 + (blk / 9) * 64, 0);
                if (++slot == Cfg::NB) { slot = 0; b_phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (convergent warp, one elected lane) =================
        constexpr uint32_t idesc = make_idesc_bf16(128, COUT);
        const uint64_t da_base = make_smem_desc(smem_u32(s_a), Cfg::SBO_A, kLayoutSw128);
        const uint64_t db_base = make_smem_desc(smem_u32(s_b), Cfg::SBO_B, kLayoutSw128);
        int slot = 0, acc = 0, it = 0;
        uint32_t b_phase = 0, acc_phase = 0;
        const uint32_t elected = elect_one() ? 1u : 0u;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            mbar_wait(&a_full[buf], ((uint32_t)it >> 1) & 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * COUT);
#pragma unroll 1
            for (int chunk = 0; chunk < 2; ++chunk) {
                const uint64_t da_c = da_base + (uint64_t)(((buf * 2 + chunk) * Cfg::CHUNK_BYTES) >> 4);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    mbar_wait(&b_full[slot], b_phase);
                    tc_fence_after();
                    const uint64_t db_s = db_base + (uint64_t)((slot * Cfg::B_BLOCK) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_pred(tmem_d, da_c + (uint64_t)((((tap / 3) * 10 + (tap % 3)) * 128 + k * 32) >> 4), db_s + (uint64_t)(k * 2), idesc,
                                       (chunk | tap | k) != 0, elected);
                    umma_commit_pred(&b_empty[slot], elected);
                    if (++slot == Cfg::NB) { slot = 0; b_phase ^= 1; }
                }
            }
            umma_commit_pred(&a_empty[buf], elected);
            umma_commit_pred(&tfull_bar[acc], elected);
            if (++acc == Cfg::NACC) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        halo_epilogue<COUT, Cfg::NACC, (Cfg::OUT_STAGE_BYTES > 0)>(p, warp, lane, tmem_base, tfull_bar, tempty_bar, s_bias, s_ostage, total_tiles, tiles_per_img);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------
// Level 1 of the UNet with a BINARY input (form_BEV = 2, every shipped configuration: learner/configs/*.txt:39-47;
// learner_models.py:489-491 turns the frame into a 0/1 mask): unet_e11 (1 -> 32, 3x3 valid, +bias, ReLU) has only 2^9
// possible outputs per pixel, one per 3x3 pattern of mask bits. The stem therefore never runs as a convolution and its
// output never exists in HBM: a tiny kernel turns the mask into one 9-bit pattern per e11 pixel, this kernel builds the
// 512 x 32-channel table of e11 values once per CTA (bf16-rounded weights, fp32 sums, bias, ReLU, bf16 -- what the
// tensor-core stem computes) and its producer warps assemble the 18x10 halo of every e12 tile by copying table rows
// into the swizzled operand layout a TMA load of e11 would have produced. MMA issue and epilogue (+ fused MaxPool2d(2))
// are those of k_tc_conv3x3_halo<32,32>. Against round 1 (stem kernel: 64 B written per pixel; e12: the same 64 B read
// back) level 1 loses 128 of its 212 bytes of HBM traffic per pixel.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_stem_patterns(const float* __restrict__ mask, uint16_t* __restrict__ pat, int N, int H, int W) {
    const int EH = H - 2, EW = W - 2;
    const long long total = (long long)N * EH * EW;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int ex = (int)(i % EW), ey = (int)((i / EW) % EH);
        const long long n = i / ((long long)EW * EH);
        const float* src = mask + (n * H + ey) * (long long)W + ex;
        unsigned bits = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) bits |= (__ldg(src + (k / 3) * W + (k % 3)) != 0.f ? 1u : 0u) << k;
        pat[i] = (uint16_t)bits;
    }
}

// form_input (learner_models.py:476-491, form_BEV = 2) and the pattern extraction in ONE pass over the frame: values below the
// cutoff are zeroed IN PLACE (the reference mutates the caller's tensor, :477) and the 0/1 mask exists only as the 9 bits
// under each e11 pixel. One thread per 8 consecutive e11 pixels of a row: 3 x 10 loads for 8 patterns. A pixel another
// thread zeroes concurrently reads as v (|v| < cutoff) or 0 -- the same mask bit either way.
__global__ void __launch_bounds__(256)
k_form_patterns(float* __restrict__ frames, float cutoff, uint16_t* __restrict__ pat, int N, int H, int W) {
    const int EH = H - 2, EW = W - 2, segs = (EW + 7) / 8;
    const long long total = (long long)N * EH * segs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int sg = (int)(i % segs), ey = (int)((i / segs) % EH);
        const long long n = i / ((long long)segs * EH);
        const int x0 = sg * 8;
        float* base = frames + (n * H + ey) * (long long)W + x0;
        unsigned rowbits[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            unsigned bits = 0;
            // this thread owns (zeroes) row ey of its columns; the last row of threads also owns the two rows below, the last
            // segment also the two columns past its patterns
            const bool own_row = r == 0 || ey == EH - 1;
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                if (x0 + j < W) {
                    float* q = base + (long long)r * W + j;
                    const float v = *q;
                    const bool small = fabsf(v) < cutoff;          // NaN compares false and stays (F8b)
                    if (small && v != 0.f && own_row && (j < 8 || sg == segs - 1)) *q = 0.f;
                    bits |= ((!small && v != 0.f) ? 1u : 0u) << j;
                }
            }
            rowbits[r] = bits;
        }
        uint16_t* o = pat + (n * EH + ey) * (long long)EW + x0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (x0 + j < EW)
                o[j] = (uint16_t)(((rowbits[0] >> j) & 7u) | (((rowbits[1] >> j) & 7u) << 3) | (((rowbits[2] >> j) & 7u) << 6));
    }
}

struct StemE12Args {
    const uint16_t* pat;      // [N, Hp-2, Wp-2] 3x3 mask patterns of the e11 pixels
    const float* stem_w;      // unet_e11.weight [32][9]
    const float* stem_b;      // [32]
};

constexpr int kStemProducerWarps = 3;   // warps 0, 2, 3: each assembles every third tile of the CTA on its own
constexpr int kLutRowB = 64;            // one table row = 32 bf16; the 16-byte chunks of a row are read in a lane-rotated order, which
                                        // spreads a warp's reads over the banks (same chunk from every lane: 16-way conflicts, ncu r2_stem_e12)

__global__ void __launch_bounds__(384, 2)
k_tc_stem_e12(const __grid_constant__ CUtensorMap map_w, const HaloArgs p, const StemE12Args sa) {
    using Cfg = HaloCfg<32, 32>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;                                  // [9][32][32] swizzled per tap
    uint8_t* s_halo = smem + Cfg::W_BYTES;                // [STAGES][180][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_halo + Cfg::STAGES * Cfg::HALO_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::STAGES;
    uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
    uint64_t* tempty_bar = tfull_bar + Cfg::NACC;
    uint64_t* w_bar = tempty_bar + Cfg::NACC;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_bar + 1);
    float* s_bias = reinterpret_cast<float*>(bars + 32);  // 16-byte aligned: the epilogue reads it as float4
    uint8_t* s_ostage = s_halo + Cfg::STAGES * Cfg::HALO_BYTES + 2048;
    uint8_t* s_lut = s_ostage + Cfg::OUT_STAGE_BYTES;     // [512] rows of 32 bf16 (64 B)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int total_tiles = tiles_per_img * p.N;
    if (warp == 0 && lane == 0) tma_prefetch_desc(&map_w);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);                   // the one producer warp that assembled the stage
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < Cfg::NACC; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 4);
        }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    if (threadIdx.x < 32) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
    // the e11 table: entry (pattern, channel pair)
    for (int i = threadIdx.x; i < 512 * 16; i += 384) {
        const int pt = i >> 4, c2 = (i & 15) * 2;
        float a0 = sa.stem_b[c2], a1 = sa.stem_b[c2 + 1];
#pragma unroll
        for (int k = 0; k < 9; ++k)
            if (pt & (1 << k)) {
                a0 += __bfloat162float(__float2bfloat16_rn(sa.stem_w[c2 * 9 + k]));
                a1 += __bfloat162float(__float2bfloat16_rn(sa.stem_w[(c2 + 1) * 9 + k]));
            }
        __nv_bfloat162 v = __floats2bfloat162_rn(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
        *reinterpret_cast<__nv_bfloat162*>(s_lut + pt * kLutRowB + (i & 15) * 4) = v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0 || warp == 2 || warp == 3) {
        // ================= halo producers: pattern -> table row -> swizzled operand row =================
        // Producer warp j assembles local tiles j, j + 3, ... on its own (180 halo pixels = 6 passes of 32 lanes), so the
        // per-tile bookkeeping runs once per tile instead of once per warp and a warp has three tile times per halo.
        const int pw = warp == 0 ? 0 : warp - 1;          // 0..2
        if (pw == 0 && lane == 0) {
            mbar_expect_tx(w_bar, Cfg::W_BYTES);
            for (int t = 0; t < 9; ++t) tma_load_2d(s_w + t * Cfg::W_TAP_BYTES, &map_w, w_bar, t * 32, 0);
        }
        const int EH = p.Hp - 2, EW = p.Wp - 2;
        const uint32_t halo_u32 = smem_u32(s_halo), lut_u32 = smem_u32(s_lut);
        constexpr int kPasses = (Cfg::HALO_ROWS + 31) / 32;                    // 6
        int hyx[kPasses];                                                      // halo pixel of (pass, lane): hy << 8 | hx
#pragma unroll
        for (int k = 0; k < kPasses; ++k) {
            const int idx = k * 32 + lane, hy = idx / 10;
            hyx[k] = idx < Cfg::HALO_ROWS ? (hy << 8) | (idx - hy * 10) : -1;
        }
        const long long first = (long long)blockIdx.x + (long long)pw * gridDim.x, step = (long long)kStemProducerWarps * gridDim.x;
        TileWalk tw(first, step, tiles_per_img);
        unsigned q[kPasses];
        auto fetch = [&]() {       // the patterns of the tile tw points at (2-byte loads from HBM, ~1 us: issued one tile of this warp ahead)
            int ty, tx;
            tw.yx(p, ty, tx);
            const uint16_t* base = sa.pat + (long long)tw.n * EH * EW;
#pragma unroll
            for (int k = 0; k < kPasses; ++k) {
                const int y = ty * 16 + (hyx[k] >> 8), x = tx * 8 + (hyx[k] & 255);
                q[k] = (hyx[k] >= 0 && y < EH && x < EW) ? (unsigned)__ldg(base + (long long)y * EW + x) : 0u;
            }
        };
        if (first < total_tiles) fetch();
        int it = pw;
        for (long long tile = first; tile < total_tiles; tile += step, it += kStemProducerWarps) {
            const int stage = it % Cfg::STAGES;
            const uint32_t phase = (uint32_t)(it / Cfg::STAGES) & 1u;
            unsigned cur[kPasses];
#pragma unroll
            for (int k = 0; k < kPasses; ++k) cur[k] = q[k];
            tw.next();
            if (tile + step < total_tiles) fetch();
            mbar_wait(&empty_bar[stage], phase ^ 1);
            const uint32_t dst = halo_u32 + stage * Cfg::HALO_BYTES;
#pragma unroll
            for (int k = 0; k < kPasses; ++k) {
                const int idx = k * 32 + lane;
                if (idx < Cfg::HALO_ROWS) {
                    const uint32_t src = lut_u32 + cur[k] * kLutRowB;
                    const uint32_t sw = (uint32_t)(idx >> 1) & 3u;
                    uint4 v[4];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) v[cc] = lds128(src + ((((uint32_t)(cc + lane)) & 3u) << 4));
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) sts128(dst + idx * 64 + (((((uint32_t)(cc + lane)) & 3u) ^ sw) << 4), v[cc]);
                }
            }
            fence_proxy_async();                                                   // this thread's generic-proxy writes -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[stage]);
        }
    } else if (warp == 1) {
        // ================= MMA issuer (as k_tc_conv3x3_halo) =================
        constexpr uint32_t idesc = make_idesc_bf16(128, 32);
        mbar_wait(w_bar, 0);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        const uint64_t dw0 = make_smem_desc(smem_u32(s_w), Cfg::SBO_B, Cfg::LAYOUT);
        const uint64_t dh0 = make_smem_desc(smem_u32(s_halo), Cfg::SBO_A, Cfg::LAYOUT);
        const uint32_t elected = elect_one() ? 1u : 0u;      // one lane issues every MMA and commit of this CTA
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 32);
            const uint64_t da0 = dh0 + (uint64_t)(stage * (Cfg::HALO_BYTES >> 4));
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint64_t da = da0 + (uint64_t)((((tap / 3) * 10 + (tap % 3)) * Cfg::ROW_B + k * 32) >> 4);
                    const uint64_t db = dw0 + (uint64_t)((tap * Cfg::W_TAP_BYTES + k * 32) >> 4);
                    umma_bf16_pred(tmem_d, da, db, idesc, (tap | k) != 0, elected);
                }
            }
            umma_commit_pred(&empty_bar[stage], elected);
            umma_commit_pred(&tfull_bar[acc], elected);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            if (++acc == Cfg::NACC) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        halo_epilogue<32, Cfg::NACC, true>(p, warp, lane, tmem_base, tfull_bar, tempty_bar, s_bias, s_ostage, total_tiles, tiles_per_img);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// input grid [N, Hp, Wp, C] as a 4-D tensor (C fastest), box [1, 18, 10, C]
static int make_map_halo(CUtensorMap* map, const void* base, int N, int Hp, int Wp, int C, int box_c = 0) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return EVFLY_ERR_CUDA;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wp * C * 2, (cuuint64_t)Hp * Wp * C * 2};
    if (box_c == 0) box_c = C;
    cuuint32_t box[4] = {(cuuint32_t)box_c, 10, 18, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle sw = (box_c * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (halo) failed with CUresult %d (N=%d Hp=%d Wp=%d C=%d)", (int)r, N, Hp, Wp, C);
        return EVFLY_ERR_CUDA;
    }
    return EVFLY_OK;
}

static int make_map_w(CUtensorMap* map, const void* base, int Cout, int Cin, int box_k = 0) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return EVFLY_ERR_CUDA;
    cuuint64_t dims[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)9 * Cin * 2};
    if (box_k == 0) box_k = Cin;
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)Cout};
    cuuint32_t es[2] = {1, 1};
    const CUtensorMapSwizzle sw = (box_k * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (halo weights) failed with CUresult %d", (int)r);
        return EVFLY_ERR_CUDA;
    }
    return EVFLY_OK;
}

template <int CIN, int COUT>
static int launch_halo(const void* x, const void* w, const HaloArgs& p, cudaStream_t st) {
    using Cfg = HaloCfg<CIN, COUT>;
    CUtensorMap mx, mw;
    int rc = make_map_halo(&mx, x, p.N, p.Hp, p.Wp, CIN);
    if (rc) return rc;
    rc = make_map_w(&mw, w, COUT, CIN);
    if (rc) return rc;
    EVFLY_SMEM_ATTR(Cfg::SMEM_BYTES, k_tc_conv3x3_halo<CIN, COUT>);
    const long long tiles = (long long)p.tiles_x * p.tiles_y * p.N;
    const int grid = (int)(tiles < Cfg::MINB * kNumSMs ? tiles : Cfg::MINB * kNumSMs);
    k_tc_conv3x3_halo<CIN, COUT><<<grid, 384, Cfg::SMEM_BYTES, st>>>(mx, mw, p);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

template <int COUT>
static int launch_halo_ws(const void* x, const void* w, const HaloArgs& p, cudaStream_t st) {
    using Cfg = HaloWsCfg<COUT>;
    CUtensorMap mx, mw;
    int rc = make_map_halo(&mx, x, p.N, p.Hp, p.Wp, 128, 64);
    if (rc) return rc;
    rc = make_map_w(&mw, w, COUT, 128, 64);
    if (rc) return rc;
    EVFLY_SMEM_ATTR(Cfg::SMEM_BYTES, k_tc_conv3x3_halo_ws<COUT>);
    const long long tiles = (long long)p.tiles_x * p.tiles_y * p.N;
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    k_tc_conv3x3_halo_ws<COUT><<<grid, 384, Cfg::SMEM_BYTES, st>>>(mx, mw, p);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

// rows of an `in`-row tensor that evfly_resize_bilinear_nhwc_bf16 (align_corners = False) reads when resizing to `out` rows. The
// kernel computes its source row in fp32 (possibly contracted to an FMA); a row within 1e-3 of a boundary is marked on both sides.
static void set_skip_rows(HaloArgs& p, int in, int out) {
    p.use_rows = 0;
    for (int i = 0; i < 16; ++i) p.row_mask[i] = 0u;
    if (out <= 0 || in > 512 || in < 1) return;
    auto mark = [&](int r) { if (r >= 0 && r < in) p.row_mask[r >> 5] |= 1u << (r & 31); };
    const double scale = (double)in / (double)out;
    for (int d = 0; d < out; ++d) {
        double src = scale * (d + 0.5) - 0.5;
        if (src < 0) src = 0;
        const int i0 = (int)src;
        mark(i0 < in - 1 ? i0 : in - 1);
        mark(i0 + 1 < in - 1 ? i0 + 1 : in - 1);
        if (src - i0 < 1e-3) mark(i0 - 1);
        if (i0 + 1 - src < 1e-3) mark(i0 + 2 < in - 1 ? i0 + 2 : in - 1);
    }
    p.use_rows = 1;
}

static void set_tiles(HaloArgs& p) {
    p.tiles_x = (p.out_vw + 7) / 8;
    p.tiles_y = (p.out_vh + 15) / 16;
    p.magic_x = (p.tiles_x > 1 && (long long)p.tiles_x * p.tiles_y < 65536) ? (uint32_t)((1ull << 32) / (unsigned)p.tiles_x) + 1u : 0u;
}

}  // namespace evfly

using namespace evfly;

static int halo_conv(const void* d_x, const void* d_w, const float* d_bias, void* d_out, void* d_pool, int N, int Hp, int Wp,
                     int vh, int vw, int Cin, int Cout, int relu, int Hp2, int Wp2, void* stream, int pad = 0, const float* d_w1 = nullptr,
                     const float* d_b1 = nullptr, float* d_out1 = nullptr, int skip_OH = 0, int compact = 0) {
    EVFLY_REQUIRE(d_x && d_w && (d_out || d_out1) && N > 0 && Hp >= 3 && Wp >= 3 && vh >= 3 && vw >= 3 && vh <= Hp && vw <= Wp, "tc_conv3x3_halo_bf16: bad shape");
    EVFLY_REQUIRE(((Cin == 32 || Cin == 64) && (Cout == 32 || Cout == 64)) || (Cin == 64 && Cout == 128) || (Cin == 128 && (Cout == 64 || Cout == 128 || Cout == 256)),
                  "tc_conv3x3_halo_bf16: (Cin, Cout) must be in {32,64}x{32,64}, (64,128) or (128, 64|128|256) (got %d, %d)", Cin, Cout);
    HaloArgs p;
    p.bias = d_bias;
    p.out = reinterpret_cast<__nv_bfloat16*>(d_out);
    p.N = N;
    p.Hp = Hp;
    p.Wp = Wp;
    p.pad = pad;
    p.out_vh = vh - 2 + 2 * pad;
    p.out_vw = vw - 2 + 2 * pad;
    p.oHp = compact ? p.out_vh : Hp;
    p.oWp = compact ? p.out_vw : Wp;
    set_tiles(p);
    p.relu = relu;
    p.pool_out = reinterpret_cast<__nv_bfloat16*>(d_pool);
    p.Hp2 = Hp2;
    p.Wp2 = Wp2;
    p.w1 = d_w1;
    p.b1 = d_b1;
    p.out1 = d_out1;
    set_skip_rows(p, p.out_vh, skip_OH);
    EVFLY_REQUIRE((d_w1 == nullptr) == (d_out1 == nullptr) && (d_w1 == nullptr) == (d_b1 == nullptr), "tc_conv3x3_halo_out1_bf16: w1 / b1 / out1 go together");
    EVFLY_REQUIRE(d_out || !d_pool, "tc_conv3x3_halo_bf16: the fused pool needs the conv output");
    EVFLY_REQUIRE(!d_pool || (Hp2 >= (vh - 2) / 2 && Wp2 >= (vw - 2) / 2), "tc_conv3x3_halo_pool_bf16: pooled grid smaller than (vh-2)/2 x (vw-2)/2");
    EVFLY_REQUIRE((long long)p.tiles_x * p.tiles_y * N < (1ll << 31), "tc_conv3x3_halo_bf16: too many tiles");
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 128 && Cout == 64) return launch_halo_ws<64>(d_x, d_w, p, st);
    if (Cin == 128 && Cout == 128) return launch_halo_ws<128>(d_x, d_w, p, st);
    if (Cin == 128 && Cout == 256) return launch_halo_ws<256>(d_x, d_w, p, st);
    if (Cin == 64 && Cout == 128) return launch_halo<64, 128>(d_x, d_w, p, st);
    if (Cin == 32 && Cout == 32) return launch_halo<32, 32>(d_x, d_w, p, st);
    if (Cin == 32 && Cout == 64) return launch_halo<32, 64>(d_x, d_w, p, st);
    if (Cin == 64 && Cout == 32) return launch_halo<64, 32>(d_x, d_w, p, st);
    return launch_halo<64, 64>(d_x, d_w, p, st);
}

extern "C" int evfly_tc_conv3x3_halo_bf16(const void* d_x, const void* d_w, const float* d_bias, void* d_out, int N, int Hp, int Wp,
                                          int vh, int vw, int Cin, int Cout, int relu, void* stream) {
    return halo_conv(d_x, d_w, d_bias, d_out, nullptr, N, Hp, Wp, vh, vw, Cin, Cout, relu, 0, 0, stream);
}

extern "C" int evfly_tc_conv3x3_halo_compact_bf16(const void* d_x, const void* d_w, const float* d_bias, void* d_out, int N, int Hp, int Wp,
                                                  int vh, int vw, int Cin, int Cout, int relu, void* stream) {
    return halo_conv(d_x, d_w, d_bias, d_out, nullptr, N, Hp, Wp, vh, vw, Cin, Cout, relu, 0, 0, stream, 0, nullptr, nullptr, nullptr, 0, 1);
}

extern "C" int evfly_tc_conv3x3_halo_pool_bf16(const void* d_x, const void* d_w, const float* d_bias, void* d_out, void* d_pool, int N, int Hp,
                                               int Wp, int vh, int vw, int Cin, int Cout, int relu, int Hp2, int Wp2, void* stream) {
    EVFLY_REQUIRE(d_pool, "tc_conv3x3_halo_pool_bf16: null pool output");
    return halo_conv(d_x, d_w, d_bias, d_out, d_pool, N, Hp, Wp, vh, vw, Cin, Cout, relu, Hp2, Wp2, stream);
}

extern "C" int evfly_tc_conv3x3_halo_pool_rows_bf16(const void* d_x, const void* d_w, const float* d_bias, void* d_out, void* d_pool, int N, int Hp,
                                                    int Wp, int vh, int vw, int Cin, int Cout, int relu, int Hp2, int Wp2, int skip_OH, void* stream) {
    EVFLY_REQUIRE(d_pool && skip_OH >= 1, "tc_conv3x3_halo_pool_rows_bf16: needs the pool output and skip_OH >= 1");
    return halo_conv(d_x, d_w, d_bias, d_out, d_pool, N, Hp, Wp, vh, vw, Cin, Cout, relu, Hp2, Wp2, stream, 0, nullptr, nullptr, nullptr, skip_OH);
}

extern "C" int evfly_tc_conv3x3_halo_out1_bf16(const void* d_x, const void* d_w, const float* d_bias, const float* d_w1, const float* d_b1, float* d_out1,
                                               int N, int Hp, int Wp, int vh, int vw, int Cin, int Cout, int relu, void* stream) {
    EVFLY_REQUIRE(d_w1 && d_b1 && d_out1, "tc_conv3x3_halo_out1_bf16: null pointer");
    EVFLY_REQUIRE((Cin == 32 || Cin == 64) && (Cout == 32 || Cout == 64), "tc_conv3x3_halo_out1_bf16: Cin, Cout in {32, 64}");
    return halo_conv(d_x, d_w, d_bias, nullptr, nullptr, N, Hp, Wp, vh, vw, Cin, Cout, relu, 0, 0, stream, 0, d_w1, d_b1, d_out1);
}

extern "C" int evfly_tc_conv3x3_same_bf16(const void* d_x, const void* d_w, const float* d_bias, void* d_out, int N, int H, int W, int Cin,
                                          int Cout, int relu, void* stream) {
    return halo_conv(d_x, d_w, d_bias, d_out, nullptr, N, H, W, H, W, Cin, Cout, relu, 0, 0, stream, 1);
}

extern "C" int evfly_stem_patterns(const float* d_mask, uint16_t* d_pat, int N, int H, int W, void* stream) {
    EVFLY_REQUIRE(d_mask && d_pat && N >= 0 && H >= 3 && W >= 3, "stem_patterns: bad argument");
    if (N == 0) return EVFLY_OK;
    const long long total = (long long)N * (H - 2) * (W - 2);
    k_stem_patterns<<<stream_grid(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(d_mask, d_pat, N, H, W);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_form_patterns(float* d_frames, float cutoff, uint16_t* d_pat, int N, int H, int W, void* stream) {
    EVFLY_REQUIRE(d_frames && d_pat && N >= 0 && H >= 3 && W >= 3, "form_patterns: bad argument");
    if (N == 0) return EVFLY_OK;
    const long long total = (long long)N * (H - 2) * ((W - 2 + 7) / 8);
    k_form_patterns<<<stream_grid(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(d_frames, cutoff, d_pat, N, H, W);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

static int stem_e12(const uint16_t* d_pat, const float* d_stem_w, const float* d_stem_b, const void* d_w, const float* d_bias, void* d_out, void* d_pool,
                    int N, int H, int W, int relu, int Hp2, int Wp2, int skip_OH, void* stream);

extern "C" int evfly_tc_stem_e12_pool_bf16(const uint16_t* d_pat, const float* d_stem_w, const float* d_stem_b, const void* d_w, const float* d_bias,
                                           void* d_out, void* d_pool, int N, int H, int W, int relu, int Hp2, int Wp2, void* stream) {
    return stem_e12(d_pat, d_stem_w, d_stem_b, d_w, d_bias, d_out, d_pool, N, H, W, relu, Hp2, Wp2, 0, stream);
}

extern "C" int evfly_tc_stem_e12_pool_rows_bf16(const uint16_t* d_pat, const float* d_stem_w, const float* d_stem_b, const void* d_w, const float* d_bias,
                                                void* d_out, void* d_pool, int N, int H, int W, int relu, int Hp2, int Wp2, int skip_OH, void* stream) {
    EVFLY_REQUIRE(skip_OH >= 1, "tc_stem_e12_pool_rows_bf16: skip_OH >= 1");
    return stem_e12(d_pat, d_stem_w, d_stem_b, d_w, d_bias, d_out, d_pool, N, H, W, relu, Hp2, Wp2, skip_OH, stream);
}

static int stem_e12(const uint16_t* d_pat, const float* d_stem_w, const float* d_stem_b, const void* d_w, const float* d_bias, void* d_out, void* d_pool,
                    int N, int H, int W, int relu, int Hp2, int Wp2, int skip_OH, void* stream) {
    EVFLY_REQUIRE(d_pat && d_stem_w && d_stem_b && d_w && d_out && N > 0 && H >= 5 && W >= 5, "tc_stem_e12_pool_bf16: bad argument");
    HaloArgs p;
    p.bias = d_bias;
    p.out = reinterpret_cast<__nv_bfloat16*>(d_out);
    p.N = N;
    p.Hp = H;
    p.Wp = W;
    p.pad = 0;
    p.out_vh = H - 4;
    p.out_vw = W - 4;
    p.oHp = H;
    p.oWp = W;
    set_tiles(p);
    p.relu = relu;
    p.pool_out = reinterpret_cast<__nv_bfloat16*>(d_pool);
    p.Hp2 = Hp2;
    p.Wp2 = Wp2;
    EVFLY_REQUIRE(!d_pool || (Hp2 >= (H - 4) / 2 && Wp2 >= (W - 4) / 2), "tc_stem_e12_pool_bf16: pooled grid smaller than (H-4)/2 x (W-4)/2");
    EVFLY_REQUIRE((long long)p.tiles_x * p.tiles_y * N < (1ll << 31), "tc_stem_e12_pool_bf16: too many tiles");
    p.w1 = nullptr;
    p.b1 = nullptr;
    p.out1 = nullptr;
    set_skip_rows(p, p.out_vh, skip_OH);
    StemE12Args sa;
    sa.pat = d_pat;
    sa.stem_w = d_stem_w;
    sa.stem_b = d_stem_b;
    using Cfg = HaloCfg<32, 32>;
    constexpr int smem = Cfg::SMEM_BYTES + 512 * kLutRowB;
    CUtensorMap mw;
    const int rc = make_map_w(&mw, d_w, 32, 32);
    if (rc) return rc;
    static_assert(2 * (smem + 1024) <= 227 * 1024, "stem+e12: two CTAs per SM");
    EVFLY_SMEM_ATTR(smem, k_tc_stem_e12);
    const long long tiles = (long long)p.tiles_x * p.tiles_y * N;
    const int grid = (int)(tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs);
    k_tc_stem_e12<<<grid, 384, smem, (cudaStream_t)stream>>>(mw, p, sa);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}
