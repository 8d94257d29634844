// stages.cu -- STAGE-LEVEL entry points of the C ABI: whole stages of the hot path enqueued from C++
// on the caller's stream, so a host in any language (the C++ side of evfly_ros, a C test) drives the path with a
// handful of calls and the whole step is capturable in one CUDA graph:
//
//   evfly_prep_frame      u8 node frame or int32 count frames -> decode, centre crop, 97th-percentile scale, clip
//                         (evfly_ros/run.py:334-350, 250-253)
//   evfly_unet_forward    OrigUNet.forward in the shipped configuration (learner/learner_models.py:521-585 with
//                         form_BEV = 2, skip_type = interp, one 1x1 ConvLSTM layer; learner/configs/*.txt:39-47):
//                         normalised frames -> depth map + y_upconv + ConvLSTM state
//   evfly_vit_lstm_forward  LSTMNetVIT.forward (learner/vitfly_models.py:132-150): depth -> velocity commands + LSTM state
//
// Weights arrive PACKED (the layouts of evfly_b200/tc.py's pack_* helpers, restated in include/evfly_b200.h) in
// caller-owned device memory; every intermediate lives in a caller-owned workspace (evfly_*_workspace_bytes);
// nothing is allocated, nothing synchronises. The functions below only SEQUENCE the operator entry points of this
// library -- the arithmetic is exactly the one of the Python drop-in modules, which call these when the model is in
// the shipped configuration.
#include "common.cuh"
#include <cuda_bf16.h>
#include <string.h>

namespace evfly {

struct Bump {
    uint8_t* p;
    int64_t left;
    bool ok = true;
    void* take(int64_t bytes) {
        const int64_t b = (bytes + 255) / 256 * 256;
        if (b > left) {
            ok = false;
            return nullptr;
        }
        void* r = p;
        p += b;
        left -= b;
        return r;
    }
};

// fp32 state [n, Ch, vh, vw] (NCHW) -> fp32 pitch grid [n, Hp, Wp, Ch] (zero outside the valid region) and back
__global__ void __launch_bounds__(256)
k_state_nchw_to_grid(const float* __restrict__ src, float* __restrict__ dst, int n, int Ch, int vh, int vw, int Hp, int Wp) {
    const long long total = (long long)n * Hp * Wp * Ch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Ch);
        const long long pix = i / Ch;
        const int x = (int)(pix % Wp), y = (int)((pix / Wp) % Hp);
        const long long b = pix / ((long long)Wp * Hp);
        dst[i] = (y < vh && x < vw) ? src[((b * Ch + c) * vh + y) * vw + x] : 0.f;
    }
}
__global__ void __launch_bounds__(256)
k_state_grid_to_nchw(const float* __restrict__ src, float* __restrict__ dst, int n, int Ch, int vh, int vw, int Hp, int Wp) {
    const long long total = (long long)n * Ch * vh * vw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % vw), y = (int)((i / vw) % vh);
        const int c = (int)((i / ((long long)vw * vh)) % Ch);
        const long long b = i / ((long long)vw * vh * Ch);
        dst[i] = src[((b * Hp + y) * Wp + x) * (long long)Ch + c];
    }
}

struct UNetGeom {
    int Hp[5], Wp[5];      // pitch grid of level l (level 0 = the input frame)
    int ev[5][2];          // valid extent of the level's second conv output (y_e{l+1})
    bool ok;
};

static UNetGeom unet_geometry(int H, int W) {
    UNetGeom g;
    g.ok = true;
    int h = H, w = W;
    for (int l = 0; l < 5; ++l) {
        g.Hp[l] = h;
        g.Wp[l] = w;
        g.ev[l][0] = h - 4;
        g.ev[l][1] = w - 4;
        if (h - 4 < 2 || w - 4 < 2) g.ok = false;
        h = (h - 4) / 2;
        w = (w - 4) / 2;
    }
    return g;
}

static inline bool halo_ok(int cin, int cout) {
    return ((cin == 32 || cin == 64) && (cout == 32 || cout == 64)) || (cin == 64 && cout == 128) || (cin == 128 && (cout == 64 || cout == 128 || cout == 256));
}

// 3x3 valid conv + bias + ReLU on the pitch grid (+ optional MaxPool2d(2) of the result), the dispatch of tc.conv3x3(_pool).
// *oHp x *oWp: pitch of `out` -- its own valid extent (compact grids: a consumer that computes every row of its grid -- the generic
// kernel, the transposed convs, the ConvLSTM -- then has no don't-care rows), the input's for the fused-pool convs (the skip reads those).
static int conv3(const void* x, int N, int Hp, int Wp, int vh, int vw, int Cin, const void* w, const float* b, int Cout, void* out, void* pool,
                 int Hp2, int Wp2, void* st, int* oHp, int* oWp, int skip_OH = 0) {
    if (halo_ok(Cin, Cout)) {
        *oHp = pool ? Hp : vh - 2;
        *oWp = pool ? Wp : vw - 2;
        if (pool && skip_OH > 0) return evfly_tc_conv3x3_halo_pool_rows_bf16(x, w, b, out, pool, N, Hp, Wp, vh, vw, Cin, Cout, 1, Hp2, Wp2, skip_OH, st);
        if (pool) return evfly_tc_conv3x3_halo_pool_bf16(x, w, b, out, pool, N, Hp, Wp, vh, vw, Cin, Cout, 1, Hp2, Wp2, st);
        return evfly_tc_conv3x3_halo_compact_bf16(x, w, b, out, N, Hp, Wp, vh, vw, Cin, Cout, 1, st);
    }
    evfly_tc_conv_args a;
    memset(&a, 0, sizeof(a));
    a.x = x;
    a.w = w;
    a.bias = b;
    a.out = out;
    a.M_rows = (int64_t)N * Hp * Wp;
    a.out_ld = Cout;
    a.Cin = Cin;
    a.n_rows = Cout;
    a.taps = 9;
    a.w_pitch = Wp;
    a.relu = 1;
    a.flags = EVFLY_TC_COMPACT;
    a.Hp = Hp;
    a.Wp = Wp;
    a.valid_h = *oHp = vh - 2;
    a.valid_w = *oWp = vw - 2;
    int rc = evfly_tc_conv_bf16(&a, st);
    if (!rc && pool) rc = evfly_maxpool2x2_nhwc_bf16(out, pool, N, vh - 2, vw - 2, vh - 2, vw - 2, Cout, st);
    return rc;
}

static int gemm_f32out(const void* x, int64_t M, int K, const void* w, const float* bias, int Nn, float* out, void* st) {
    evfly_tc_conv_args a;
    memset(&a, 0, sizeof(a));
    a.x = x;
    a.w = w;
    a.bias = bias;
    a.out_f32 = out;
    a.M_rows = M;
    a.out_ld = Nn;
    a.Cin = K;
    a.n_rows = Nn;
    a.taps = 1;
    return evfly_tc_conv_bf16(&a, st);
}

static int convt2x2(const void* x, int N, int Hp, int Wp, int vh, int vw, int Cin, const void* w, const float* b, int Cout, void* out, int out_ld,
                    int out_c0, void* st) {
    evfly_tc_conv_args a;
    memset(&a, 0, sizeof(a));
    a.x = x;
    a.w = w;
    a.bias = b;
    a.out = out;
    a.M_rows = (int64_t)N * Hp * Wp;
    a.out_ld = out_ld;
    a.Cin = Cin;
    a.n_rows = 4 * Cout;
    a.taps = 1;
    a.out_c0 = out_c0;
    a.convt = 1;
    a.Hp = Hp;
    a.Wp = Wp;
    a.valid_h = vh;
    a.valid_w = vw;
    a.cout_t = Cout;
    return evfly_tc_conv_bf16(&a, st);
}

static const int kEncC[6] = {1, 32, 64, 128, 256, 512};

}  // namespace evfly

using namespace evfly;

#define RC(call)                \
    do {                        \
        const int rc__ = (call); \
        if (rc__) return rc__;  \
    } while (0)

extern "C" int evfly_prep_frame(const uint8_t* d_u8, const int32_t* d_counts, int N, int H, int W, int h, int w, float* d_frames, void* stream) {
    EVFLY_REQUIRE((d_u8 != nullptr) != (d_counts != nullptr), "prep_frame: give either the u8 node frames or the int32 count frames");
    EVFLY_REQUIRE(d_frames && N >= 0 && h <= H && w <= W, "prep_frame: bad argument");
    if (N == 0) return EVFLY_OK;
    if (d_counts)      // integer counts: decode + crop + exact percentile + clip in one kernel
        return evfly_counts_normalise(d_counts, N, H, W, h, w, 0.2f, 0.97f, -1.0f, 1.0f, 0.0f, d_frames, nullptr, stream);
    RC(evfly_decode_crop(d_u8, nullptr, N, H, W, h, w, 0.2f, d_frames, stream));
    return evfly_quantile_scale_clip(d_frames, N, (int64_t)h * w, 0.97f, -1.0f, 1.0f, 0.0f, d_frames, nullptr, stream);
}

extern "C" int64_t evfly_unet_workspace_bytes(int N, int n_traj, int H, int W) {
    const UNetGeom g = unet_geometry(H, W);
    if (!g.ok || N <= 0 || n_traj <= 0 || N % n_traj) return 0;
    int64_t total = 4096;
    auto add = [&](int64_t b) { total += (b + 255) / 256 * 256; };
    add((int64_t)N * (H - 2) * (W - 2) * 2);          // stem patterns
    for (int l = 0; l < 5; ++l) {
        const int64_t px = (int64_t)N * g.Hp[l] * g.Wp[l];
        if (l > 0) add(px * kEncC[l] * 2);            // pooled input of the level
        if (l > 0) add(px * kEncC[l + 1] * 2);        // first conv
        add(px * kEncC[l + 1] * 2);                   // second conv (y_e)
    }
    const int64_t P5 = (int64_t)N * g.Hp[4] * g.Wp[4];
    add(P5 * 2048 * 4);                               // x gates
    add((P5 + P5 / (N / n_traj)) * 512 * 2);          // h_all
    add(P5 / (N / n_traj) * 512 * 4);                 // c
    add(256);                                         // scan barrier
    int vh = g.ev[4][0], vw = g.ev[4][1];
    for (int lvl = 1; lvl <= 4; ++lvl) {
        const int C = kEncC[5 - lvl];
        const int64_t px = (int64_t)N * (2 * vh) * (2 * vw);
        add(px * 2 * C * 2);                          // cat
        add(px * C * 2);                              // d_l1
        add(px * C * 2);                              // d_l2
        vh = 2 * vh - 4;
        vw = 2 * vw - 4;
    }
    add((int64_t)N * (vh + 4) * (vw + 4) * 4);        // out32
    return total;
}

extern "C" int evfly_unet_forward(const evfly_unet_weights* wts, float* d_frames, int N, int n_traj, int H, int W, float cutoff, const float* d_h0,
                                  const float* d_c0, float* d_hT, float* d_cT, float* d_depth, float* d_y_upconv, void* d_ws, int64_t ws_bytes,
                                  void* stream) {
    EVFLY_REQUIRE(wts && d_frames && d_depth && d_y_upconv && d_ws && N > 0 && n_traj > 0 && N % n_traj == 0, "unet_forward: bad argument");
    EVFLY_REQUIRE((d_h0 == nullptr) == (d_c0 == nullptr), "unet_forward: h0 / c0 go together");
    const UNetGeom g = unet_geometry(H, W);
    EVFLY_REQUIRE(g.ok, "unet_forward: a %dx%d frame is too small for the 5-level UNet", H, W);
    const int64_t need = evfly_unet_workspace_bytes(N, n_traj, H, W);
    if (ws_bytes < need) {
        set_error("unet_forward: workspace of %lld bytes, %lld needed", (long long)ws_bytes, (long long)need);
        return EVFLY_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Bump ws{reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(d_ws) + 255) & ~(uintptr_t)255), ws_bytes - 256};
    const int T = N / n_traj;

    // ---- form_input (learner_models.py:476-494, form_BEV = 2: cutoff in place on the caller's frames, 0/1 mask) fused with the
    //      extraction of the stem's 3x3 mask patterns; ---- encoder (learner_models.py:533-541)
    void* ye[5];
    int yeHp[5] = {H, 0, 0, 0, 0}, yeWp[5] = {W, 0, 0, 0, 0};      // pitch of y_e{l+1}
    {
        uint16_t* pat = (uint16_t*)ws.take((int64_t)N * (H - 2) * (W - 2) * 2);
        RC(evfly_form_patterns(d_frames, cutoff, pat, N, H, W, stream));
        ye[0] = ws.take((int64_t)N * H * W * 32 * 2);
        void* pooled = ws.take((int64_t)N * g.Hp[1] * g.Wp[1] * 32 * 2);
        // y_e1..y_e3 are only sampled by the decoder's bilinear skip: heights 2 * (valid height of that decoder level's input)
        int skip_oh[4];
        {
            int vh = g.ev[4][0];
            for (int el = 3; el >= 0; --el) {
                skip_oh[el] = 2 * vh;
                vh = 2 * vh - 4;
            }
        }
        RC(evfly_tc_stem_e12_pool_rows_bf16(pat, wts->e11_w, wts->e11_b, wts->conv_w[0], wts->conv_b[0], ye[0], pooled, N, H, W, 1, g.Hp[1], g.Wp[1], skip_oh[0], stream));
        const void* x = pooled;
        for (int l = 1; l < 5; ++l) {
            const int Hp = g.Hp[l], Wp = g.Wp[l], Cin = kEncC[l], C = kEncC[l + 1];
            const int64_t px = (int64_t)N * Hp * Wp;
            void* c1 = ws.take(px * C * 2);
            ye[l] = ws.take(px * C * 2);
            void* nxt = l < 4 ? ws.take((int64_t)N * g.Hp[l + 1] * g.Wp[l + 1] * C * 2) : nullptr;
            int h1, w1;
            RC(conv3(x, N, Hp, Wp, Hp, Wp, Cin, wts->conv_w[2 * l - 1], wts->conv_b[2 * l - 1], C, c1, nullptr, 0, 0, stream, &h1, &w1));
            RC(conv3(c1, N, h1, w1, Hp - 2, Wp - 2, C, wts->conv_w[2 * l], wts->conv_b[2 * l], C, ye[l], nxt, l < 4 ? g.Hp[l + 1] : 0, l < 4 ? g.Wp[l + 1] : 0, stream,
                     &yeHp[l], &yeWp[l], l < 3 ? skip_oh[l] : 0));
            x = nxt;
        }
    }
    // ---- ConvLSTM over time (learner_models.py:544-546; convlstm.py:136-176), n_traj trajectories side by side
    const int Hp5 = yeHp[4], Wp5 = yeWp[4], vh5 = g.ev[4][0], vw5 = g.ev[4][1], Ch = 512;
    const int64_t P = (int64_t)n_traj * Hp5 * Wp5;
    float* gx = (float*)ws.take((int64_t)T * P * 4 * Ch * 4);
    __nv_bfloat16* h_all = (__nv_bfloat16*)ws.take((int64_t)(T + 1) * P * Ch * 2);
    float* cst = (float*)ws.take(P * Ch * 4);
    void* sync = ws.take(256);
    EVFLY_REQUIRE(ws.ok, "unet_forward: workspace exhausted (internal sizing error)");
    if (d_h0) {
        RC(evfly_nchw_f32_to_nhwc_bf16(d_h0, h_all, n_traj, Ch, vh5, vw5, Hp5, Wp5, stream));
        k_state_nchw_to_grid<<<stream_grid(P * Ch, 256, 8), 256, 0, st>>>(d_c0, cst, n_traj, Ch, vh5, vw5, Hp5, Wp5);
        EVFLY_LAUNCHED();
    } else {
        EVFLY_CUDA(cudaMemsetAsync(h_all, 0, P * Ch * 2, st));
        EVFLY_CUDA(cudaMemsetAsync(cst, 0, P * Ch * 4, st));
    }
    {   // x-gates inside the persistent step kernel; where that grid cannot be co-resident: the x-gate GEMM + the per-step scan
        const int rc_scan = evfly_convlstm_scan_fused_bf16(ye[4], wts->lstm_wx, h_all, wts->lstm_wh, cst, T, P, Ch, Ch, sync, stream);
        if (rc_scan == EVFLY_ERR_UNSUPPORTED) {
            RC(gemm_f32out(ye[4], (int64_t)T * P, Ch, wts->lstm_wx, nullptr, 4 * Ch, gx, stream));
            RC(evfly_convlstm_scan_bf16(h_all, wts->lstm_wh, gx, cst, T, P, Ch, sync, stream));
        } else {
            RC(rc_scan);
        }
    }
    if (d_hT) RC(evfly_nhwc_to_nchw_f32(h_all + (int64_t)T * P * Ch, 0, d_hT, n_traj, Ch, vh5, vw5, Hp5, Wp5, stream));
    if (d_cT) {
        k_state_grid_to_nchw<<<stream_grid((int64_t)n_traj * Ch * vh5 * vw5, 256, 8), 256, 0, st>>>(cst, d_cT, n_traj, Ch, vh5, vw5, Hp5, Wp5);
        EVFLY_LAUNCHED();
    }
    // ---- decoder (learner_models.py:553-585): bilinear skip || ConvTranspose2d(2,2) -> cat -> two 3x3 convs
    const void* y = h_all + P * Ch;       // h_1 .. h_T = the ConvLSTM output sequence, frame t*n_traj + s
    float* out32 = nullptr;
    int yHp = Hp5, yWp = Wp5, yvh = vh5, yvw = vw5;
    for (int lvl = 1; lvl <= 4; ++lvl) {
        const int C = kEncC[5 - lvl], Cin = kEncC[6 - lvl];
        const int oh = 2 * yvh, ow = 2 * yvw;
        const int64_t px = (int64_t)N * oh * ow;
        void* cat = ws.take(px * 2 * C * 2);
        void* d1 = ws.take(px * C * 2);
        void* d2 = ws.take(px * C * 2);
        EVFLY_REQUIRE(ws.ok, "unet_forward: workspace exhausted (internal sizing error)");
        const int el = 4 - lvl;           // encoder level whose output is skipped in
        RC(evfly_resize_bilinear_nhwc_bf16(ye[el], cat, N, yeHp[el], yeWp[el], g.ev[el][0], g.ev[el][1], C, oh, ow, 2 * C, 0, stream));
        RC(convt2x2(y, N, yHp, yWp, yvh, yvw, Cin, wts->up_w[lvl - 1], wts->up_b[lvl - 1], C, cat, 2 * C, C, stream));
        int h1, w1, h2 = oh, w2 = ow;
        RC(conv3(cat, N, oh, ow, oh, ow, 2 * C, wts->conv_w[9 + 2 * (lvl - 1)], wts->conv_b[9 + 2 * (lvl - 1)], C, d1, nullptr, 0, 0, stream, &h1, &w1));
        if (lvl == 4) {
            // unet_d42 with unet_out (1x1 to one channel, :583) in its epilogue: the 32-channel activation is never written
            out32 = (float*)d2;       // N*oh*ow fp32 fit the N*oh*ow*32 bf16 reserved for d42's output
            RC(evfly_tc_conv3x3_halo_out1_bf16(d1, wts->conv_w[16], wts->conv_b[16], wts->out_w, wts->out_b, out32, N, h1, w1, oh - 2, ow - 2, C, C, 1, stream));
            h2 = h1;
            w2 = w1;
        } else {
            RC(conv3(d1, N, h1, w1, oh - 2, ow - 2, C, wts->conv_w[10 + 2 * (lvl - 1)], wts->conv_b[10 + 2 * (lvl - 1)], C, d2, nullptr, 0, 0, stream, &h2, &w2));
        }
        y = d2;
        yHp = h2;
        yWp = w2;
        yvh = oh - 4;
        yvw = ow - 4;
    }
    // ---- 1x1 output conv (:583) and the bilinear resize back to the input size (:497)
    RC(evfly_nhwc_to_nchw_f32(out32, 1, d_y_upconv, N, 1, yvh, yvw, yHp, yWp, stream));
    const int64_t xs[4] = {(int64_t)yvh * yvw, (int64_t)yvh * yvw, yvw, 1}, ys[4] = {(int64_t)H * W, (int64_t)H * W, W, 1};
    return evfly_resize_bilinear_f32(d_y_upconv, xs, d_depth, ys, N, 1, yvh, yvw, H, W, 0, 1.0f, 0.0f, -INFINITY, INFINITY, stream);
}

// ======================================================================================================================
// LSTMNetVIT.forward (learner/vitfly_models.py:132-150), bf16 tensor-core path, batches (N >= 8 frames):
//   depth [N,1,H,W] -> (clamp(2 d, 0, 1) when premap) -> bilinear 60x90 -> two Mix-Transformer stages (fused kernels of
//   vit_fused.cu) -> PixelShuffle / Upsample / cat -> conv 48->12 -> Linear 4608->512 -> cat[., desvel/10, quat] ->
//   LSTM(517 -> 128, 3 layers) over time -> Linear 128 -> 3
// ======================================================================================================================
namespace evfly {

static int linear_f32(const float* x, int64_t x_ld, int64_t M, int K, const float* w, const float* bias, float* y, int64_t y_ld, int Nout, void* st) {
    if (M <= 8 && (int64_t)M * K * 4 <= 160 * 1024)
        return evfly_linear_smallm_f32(x, x_ld, w, bias, nullptr, 0, y, y_ld, (int)M, Nout, K, EVFLY_ACT_NONE, st);
    evfly_conv2d_args a;
    memset(&a, 0, sizeof(a));
    a.x = x;
    a.w = w;
    a.bias = bias;
    a.y = y;
    a.N = 1;
    a.Cin = K;
    a.H = 1;
    a.W = (int32_t)M;
    a.Cout = Nout;
    a.KH = a.KW = 1;
    a.stride = 1;
    a.groups = 1;
    a.xs[1] = 1;
    a.xs[3] = x_ld;
    a.ys[1] = 1;
    a.ys[3] = y_ld;
    return evfly_conv2d_f32(&a, st);
}

__global__ void __launch_bounds__(256)
k_seq_tail(float* __restrict__ seq, const float* __restrict__ desvel, const float* __restrict__ quat, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float* r = seq + (size_t)i * 517 + 512;
    r[0] = desvel[i] / 10.0f;                                  // X[1] / 10 (vitfly_models.py:144)
    if (quat) {
        r[1] = quat[4 * i]; r[2] = quat[4 * i + 1]; r[3] = quat[4 * i + 2]; r[4] = quat[4 * i + 3];
    } else {                                                   // refine_inputs' default quaternion [1,0,0,0] (:22-25)
        r[1] = 1.f; r[2] = 0.f; r[3] = 0.f; r[4] = 0.f;
    }
}

}  // namespace evfly

extern "C" int64_t evfly_vit_lstm_workspace_bytes(int N) {
    if (N <= 0) return 0;
    int64_t t = 4096;
    auto add = [&](int64_t b) { t += (b + 255) / 256 * 256; };
    add((int64_t)N * 60 * 90 * 4);                 // resized depth
    add((int64_t)N * 345 * 32 * 2 * 3);            // stage-1 tokens (ping, pong, attention out)
    add((int64_t)N * 96 * 64 * 2 * 3);             // stage-2 tokens
    for (int s = 0; s < 2; ++s) {                  // per stage: reduced tokens and their K/V projection
        add((int64_t)N * 6 * 64 * 2);
        add((int64_t)N * 6 * 128 * 2);
    }
    add((int64_t)N * 16 * 24 * 64 * 2);            // tail cat
    add((int64_t)N * 16 * 24 * 32 * 2);            // tail conv out
    add((int64_t)N * 517 * 4);                     // seq
    add((int64_t)N * 512 * 4);                     // gate pre-activations
    add((int64_t)N * 128 * 4 * 2);                 // layer outputs (ping, pong)
    return t;
}

extern "C" int evfly_vit_lstm_forward(const evfly_vit_lstm_weights* w, const float* d_depth, int N, int n_traj, int H, int W, int premap_clamp,
                                      const float* d_desvel, const float* d_quat, const float* d_h0, const float* d_c0, float* d_hT, float* d_cT,
                                      float* d_vel, void* d_ws, int64_t ws_bytes, void* stream) {
    EVFLY_REQUIRE(w && d_depth && d_desvel && d_vel && d_hT && d_cT && d_ws && N >= 8 && n_traj > 0 && N % n_traj == 0,
                  "vit_lstm_forward: bad argument (batches of N >= 8 frames; smaller batches go through the per-operator entry points)");
    EVFLY_REQUIRE((d_h0 == nullptr) == (d_c0 == nullptr), "vit_lstm_forward: h0 / c0 go together");
    const int64_t need = evfly_vit_lstm_workspace_bytes(N);
    if (ws_bytes < need) {
        set_error("vit_lstm_forward: workspace of %lld bytes, %lld needed", (long long)ws_bytes, (long long)need);
        return EVFLY_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Bump ws{reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(d_ws) + 255) & ~(uintptr_t)255), ws_bytes - 256};
    // ---- refine_inputs (:28-29) fused with the wrapper's clamp (learner_models.py:634)
    const float* depth60 = d_depth;
    if (H != 60 || W != 90 || premap_clamp) {
        float* r = (float*)ws.take((int64_t)N * 60 * 90 * 4);
        const int64_t xs[4] = {(int64_t)H * W, (int64_t)H * W, W, 1}, ys[4] = {60 * 90, 60 * 90, 90, 1};
        if (premap_clamp) RC(evfly_resize_bilinear_premap_f32(d_depth, xs, r, ys, N, 1, H, W, 60, 90, 0, 2.0f, 0.0f, 1.0f, stream));
        else RC(evfly_resize_bilinear_f32(d_depth, xs, r, ys, N, 1, H, W, 60, 90, 0, 1.0f, 0.0f, -INFINITY, INFINITY, stream));
        depth60 = r;
    }
    // ---- the two Mix-Transformer stages (ViTsubmodules.py:122-148)
    const void* stage_in = depth60;
    void* tok_out[2] = {nullptr, nullptr};
    const int TH[2] = {15, 8}, TW[2] = {23, 12}, TC[2] = {32, 64}, CIN[2] = {1, 32}, KP[2] = {7, 3}, SP[2] = {4, 2}, PP[2] = {3, 1}, RR[2] = {8, 4},
              HEADS[2] = {1, 2}, IH[2] = {60, 15}, IW[2] = {90, 23};
    for (int s = 0; s < 2; ++s) {
        const evfly_vit_stage_weights& sw = w->stage[s];
        const int C = TC[s], Ntok = TH[s] * TW[s];
        const int64_t tok_bytes = (int64_t)N * Ntok * C * 2;
        void* a = ws.take(tok_bytes);
        void* b = ws.take(tok_bytes);
        void* c = ws.take(tok_bytes);
        const int h2 = (TH[s] - RR[s]) / RR[s] + 1, w2 = (TW[s] - RR[s]) / RR[s] + 1, n_kv = h2 * w2;
        void* red = ws.take((int64_t)N * n_kv * C * 2);
        void* kv = ws.take((int64_t)N * n_kv * 2 * C * 2);
        EVFLY_REQUIRE(ws.ok, "vit_lstm_forward: workspace exhausted (internal sizing error)");
        RC(evfly_patch_embed_ln_bf16(stage_in, s == 0, sw.patch_w, sw.patch_b, sw.patch_ln_g, sw.patch_ln_b, a, N, IH[s], IW[s], CIN[s], C, KP[s], SP[s], PP[s],
                                     1e-5f, stream));
        void* cur = a;
        void* spare = b;
        for (int l = 0; l < 2; ++l) {
            const evfly_vit_layer_weights& lw = sw.layer[l];
            // spatial-reduction attention: k = s = r conv + LayerNorm, K/V projection, fused q / softmax / final projection + residual
            RC(evfly_patch_embed_ln_bf16(cur, 0, lw.red_w, lw.red_b, lw.red_ln_g, lw.red_ln_b, red, N, TH[s], TW[s], C, C, RR[s], RR[s], 0, 1e-5f, stream));
            evfly_tc_conv_args g;
            memset(&g, 0, sizeof(g));
            g.x = red;
            g.w = lw.kv_w;
            g.bias = lw.kv_b;
            g.out = kv;
            g.M_rows = (int64_t)N * n_kv;
            g.out_ld = 2 * C;
            g.Cin = C;
            g.n_rows = 2 * C;
            g.taps = 1;
            RC(evfly_tc_conv_bf16(&g, stream));
            RC(evfly_vit_attn_bf16(cur, kv, lw.attn_img, lw.attn_bias, c, N, Ntok, C, HEADS[s], n_kv, stream));
            // MixFFN + residual + LayerNorm in one launch
            RC(evfly_vit_ffn_bf16(c, lw.ffn_img, lw.ffn_bias, spare, N, TH[s], TW[s], C, 1e-5f, stream));
            void* t = cur;
            cur = spare;
            spare = t;
        }
        tok_out[s] = cur;
        stage_in = cur;
    }
    // ---- tail (vitfly_models.py:136-143): cat[PixelShuffle(2)(s2), Upsample(16x24, align_corners)(s1)] -> conv 48->12 -> decoder Linear
    void* cat = ws.take((int64_t)N * 16 * 24 * 64 * 2);
    void* feat = ws.take((int64_t)N * 16 * 24 * 32 * 2);
    float* seq = (float*)ws.take((int64_t)N * 517 * 4);
    float* gx = (float*)ws.take((int64_t)N * 512 * 4);
    float* hs0 = (float*)ws.take((int64_t)N * 128 * 4);
    float* hs1 = (float*)ws.take((int64_t)N * 128 * 4);
    EVFLY_REQUIRE(ws.ok, "vit_lstm_forward: workspace exhausted (internal sizing error)");
    RC(evfly_shuffle_upsample_cat_bf16(tok_out[1], 8, 12, 64, tok_out[0], 15, 23, 32, cat, N, 64, stream));
    RC(evfly_tc_conv3x3_same_bf16(cat, w->ds_w, w->ds_b, feat, N, 16, 24, 64, 32, 0, stream));
    {
        evfly_tc_conv_args g;
        memset(&g, 0, sizeof(g));
        g.x = feat;
        g.w = w->dec_w;
        g.bias = w->dec_b;
        g.out_f32 = seq;
        g.M_rows = N;
        g.out_ld = 517;
        g.Cin = 16 * 24 * 32;
        g.n_rows = 512;
        g.taps = 1;
        RC(evfly_tc_conv_bf16(&g, stream));
    }
    k_seq_tail<<<(N + 255) / 256, 256, 0, st>>>(seq, d_desvel, d_quat, N);
    EVFLY_LAUNCHED();
    // ---- LSTM over time (:146), n_traj sequences side by side, then the velocity head (:149)
    const int T = N / n_traj;
    const float* inp = seq;
    int in_ld = 517, in_k = 517;
    float* outs[2] = {hs0, hs1};
    for (int l = 0; l < 3; ++l) {
        RC(linear_f32(inp, in_ld, N, in_k, w->lstm_w_ih[l], w->lstm_b[l], gx, 512, 512, stream));
        float* hs = outs[l & 1];
        const float* h0 = d_h0 ? d_h0 + (size_t)l * n_traj * 128 : nullptr;
        const float* c0 = d_c0 ? d_c0 + (size_t)l * n_traj * 128 : nullptr;
        if (T >= 16) RC(evfly_lstm_seq_smemw(gx, w->lstm_whh_pairs[l], h0, c0, hs, d_hT + (size_t)l * n_traj * 128, d_cT + (size_t)l * n_traj * 128, T, 128, n_traj, stream));
        else RC(evfly_lstm_seq_f32(gx, w->lstm_whh_t[l], h0, c0, hs, d_hT + (size_t)l * n_traj * 128, d_cT + (size_t)l * n_traj * 128, T, 128, n_traj, stream));
        inp = hs;
        in_ld = in_k = 128;
    }
    return linear_f32(inp, 128, N, 128, w->fc2_w, w->fc2_b, d_vel, 3, 3, stream);
}
