/*
 * evfly_b200.h -- C ABI of libevfly_b200.so: the B200 (sm_100a) implementation of
 * evfly's perception hot path (event accumulation -> frame normalisation ->
 * depth-pretext UNet/ConvLSTM -> ViT-LSTM velocity forward).
 *
 * Conventions:
 *   - every pointer named d_* is CALLER-OWNED DEVICE memory (e.g. a torch tensor's
 *     data_ptr()); the library keeps no reference after the call is enqueued;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no
 *     hidden device synchronisation and no allocation, and is safe to capture in a CUDA graph;
 *   - scratch memory is caller-owned too: ask evfly_*_workspace_bytes() and pass a buffer;
 *     scratch that must be zero on entry is documented as such and is LEFT ZERO on exit;
 *   - return value 0 on success, a negative EVFLY_ERR_* otherwise; evfly_last_error()
 *     returns a thread-local human-readable message. No exception crosses the ABI;
 *   - there is no CPU fallback anywhere in the library.
 *
 * Reference interfaces replaced (paths relative to the evfly tree):
 *   evfly_ros/src/node.cpp:24-59        ImagePublisher::eventArrayCallback/timerCallback (u8, wraps)
 *   evfly_dv_ros/src/node.cpp:24-63     same for DAVIS (u8, saturates at 0/255)
 *   utils/ev_utils.py:113-161           form_eventframe (np.histogram2d signed count frame)
 *   utils/to_events.py:400-411          per-window slicing of one continuous stream
 *   evfly_ros/run.py:334-350,250-253    u8 decode, centre-crop, 97th-percentile scale + clip
 *   learner/dataloading.py:518-521,531-533  per-frame percentile scale, clamp, min-cutoff
 *   learner/learner_models.py, learner/vitfly_models.py, learner/ViTsubmodules.py,
 *   learner/ConvLSTM_pytorch/convlstm.py    model forward (see the model section below)
 */
#ifndef EVFLY_B200_H
#define EVFLY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVFLY_ABI_VERSION 1

/* ---- error codes ------------------------------------------------------------------- */
#define EVFLY_OK                 0
#define EVFLY_ERR_INVALID_ARG   -1   /* null pointer, negative size, unsupported shape ...   */
#define EVFLY_ERR_CUDA          -2   /* a CUDA runtime call failed (message has the detail) */
#define EVFLY_ERR_WORKSPACE     -3   /* workspace too small                                 */
#define EVFLY_ERR_UNSUPPORTED   -4   /* configuration not implemented                       */

/* ---- the packed event record ---------------------------------------------------------
 * 16 bytes, identical to the in-memory layout of dv_ros_msgs::Event /
 * prophesee_event_msgs::Event (dv_ros_msgs/msg/Event.msg:2-5: uint16 x, uint16 y, time ts,
 * bool polarity; ros::Time is {uint32 sec, uint32 nsec}), so a ROS node can memcpy
 * msg->events.data() straight into a pinned staging buffer. One ld.global.v4.b32 per event.
 * polarity: 1 = positive (++), 0 = negative (--), >= 2 = record is skipped (written by the
 * packers for events that the reference's masks drop; never produced by a sensor).        */
typedef struct evfly_event {
    uint16_t x;
    uint16_t y;
    uint32_t ts_sec;
    uint32_t ts_nsec;
    uint8_t  polarity;
    uint8_t  pad[3];
} evfly_event;

#define EVFLY_POL_NEG   0
#define EVFLY_POL_POS   1
#define EVFLY_POL_SKIP  2

/* polarity convention of the float [t,x,y,p] rows handed to form_eventframe            */
#define EVFLY_NEG_IS_ZERO      0   /* all_events=True : pos p>0, neg p==0 (ev_utils.py:155-156) */
#define EVFLY_NEG_IS_NEGATIVE  1   /* timed / to_events: pos p>0, neg p<0 (ev_utils.py:137-138) */

/* u8 accumulator overflow behaviour */
#define EVFLY_U8_WRAP      0   /* evfly_ros/src/node.cpp:33-37 (uint8 ++/-- wraps mod 256)   */
#define EVFLY_U8_SATURATE  1   /* evfly_dv_ros/src/node.cpp:33-41 (clamps at 0 / 255)        */

/* ---- library ------------------------------------------------------------------------ */
int         evfly_abi_version(void);
const char* evfly_last_error(void);
/* number of kernel launches this process has enqueued through the library (bench.py's
 * gpu_launches claim is read from here).                                                 */
int64_t     evfly_launch_count(void);

/* ======================================================================================
 * L1  event accumulation
 * ====================================================================================== */

/* Pack float64 rows [n,4] = (t, x, y, p) -- the argument form_eventframe receives -- into
 * evfly_event records, applying exactly the reference's masks:
 *   - np.histogram2d binning over range [0,W]x[0,H] with W x H unit bins: bin = floor(coord),
 *     the right edge is included (x == W -> bin W-1), everything else (incl. NaN) dropped;
 *   - polarity classes per `pol_mode`;
 *   - if use_time: keep only t_lo <= t < t_hi, compared in float64 like ev_utils.py:128
 *     (the caller passes times0*1e9 and times1[0]*1e9 computed in float64);
 *   - if max_events >= 0 (the `N is not None` branch, ev_utils.py:130-133): keep only the first
 *     max_events rows with t >= t_lo; *d_last_kept_t receives the t of the last kept row.
 * Dropped rows become polarity=EVFLY_POL_SKIP records, so out[i] always corresponds to row i.
 * ts is floor(t) split into sec/nsec (t is in ns, ev_utils.py:128).
 * d_scan_ws: (n/1024+2)*8 bytes of scratch, only needed when max_events >= 0.            */
int evfly_pack_events_f64(const double* d_rows, int64_t n, int H, int W, int pol_mode,
                          int use_time, double t_lo, double t_hi, int64_t max_events,
                          evfly_event* d_out, double* d_last_kept_t, void* d_scan_ws,
                          void* stream);

/* Pack structure-of-arrays streams (utils/to_events.py:400-411: events[traj]['x','y','t','p']
 * with p in {+1,-1} and t in ns) into records. Any of the integer types is given as int64. */
int evfly_pack_events_soa(const int64_t* d_x, const int64_t* d_y, const int64_t* d_t_ns,
                          const int64_t* d_p, int64_t n, int H, int W, int pol_mode,
                          evfly_event* d_out, void* stream);

/* ---- the 8-byte wire record ------------------------------------------------------------
 * The canonical record with the absolute timestamp replaced by the offset from the first edge of the window
 * the event has been assigned to (the host that slices a stream into windows -- utils/to_events.py:400-411,
 * the 30 Hz timer of evfly_ros/src/node.cpp:42-59 -- knows that window). Half the bytes per event over PCIe.
 * x == 0xffff marks a skipped record.                                                              */
typedef struct evfly_event8 {
    uint16_t x;
    uint16_t y;
    uint32_t dt_pol;      /* (t_ns - window_t0_ns) << 1 | polarity ; offsets up to 2^31 ns = 2.1 s */
} evfly_event8;

/* Canonical records of a time-sorted stream -> wire records, position by position, plus the event index range
 * of every window: d_win_offsets int64 [T+1] (window w = records [off[w], off[w+1])). Records outside
 * [edges[0], edges[T]), with polarity >= 2 or farther than 2^31 ns from their window's start become skips. */
int evfly_pack_events_ev8(const evfly_event* d_events, int64_t n, const int64_t* d_edges_ns, int T,
                          evfly_event8* d_out, int64_t* d_win_offsets, void* stream);

/* Windows of a stream whose events are grouped by window, through shared-memory histogram tiles
 * (evfly_b200/csrc/accumulate_tiled.cu): every chunk of 8192 events is counting-sorted by row band in shared
 * memory, then one CTA per (band, window) accumulates the band's 2 + B planes in shared memory and writes them
 * out once. No zero-fill, no global reduction; same results as evfly_accumulate_windows (counts bit-exact, voxel
 * sums in a different order). Windows must be shorter than 2^32 ns (events farther from their window's start
 * are dropped).
 *   evfly_accumulate_windows_sorted : canonical records of ONE time-sorted stream, edges int64 [T+1]; window w
 *       goes to frame slot w*slot_stride + slot_offset (1, 0: consecutive; n, s: trajectory s of a time-major batch)
 *   evfly_accumulate_windows_ev8    : wire records of any number of streams laid end to end; window w is
 *       records [d_win_offsets[w], d_win_offsets[w+1]) with time range [d_win_t0[w], d_win_t1[w]); its frame goes
 *       to slot d_out_slot[w] of d_counts / d_voxel (NULL: slot w) -- e.g. time-major order for a batch of
 *       trajectories (learner/evaluation_tools.py:62-66 evaluates them one by one).
 * d_ws: evfly_accumulate_sorted_workspace_bytes(n, n_windows, H, W, B) bytes of scratch (any content).      */
int64_t evfly_accumulate_sorted_workspace_bytes(int64_t n, int n_windows, int H, int W, int B);
int evfly_accumulate_windows_sorted(const evfly_event* d_events, int64_t n, const int64_t* d_edges_ns, int T,
                                    int H, int W, int B, int32_t* d_counts, float* d_voxel, int slot_stride,
                                    int slot_offset, void* d_ws, int64_t ws_bytes, void* stream);
int evfly_accumulate_windows_ev8(const evfly_event8* d_events, int64_t n, const int64_t* d_win_offsets,
                                 const int64_t* d_win_t0, const int64_t* d_win_t1, const int32_t* d_out_slot,
                                 int n_windows, int H, int W, int B, int32_t* d_counts, float* d_voxel, void* d_ws,
                                 int64_t ws_bytes, void* stream);

/* The 4-byte wire record (uint32): x [0,10) | y [10,19) | polarity [19] | delta [20,32), delta = microseconds since the previous
 * record of the same window (since the window's first edge for its first record); (x, y) = (1023, 511) is a skip record that
 * only advances the time (gaps above 4095 us). It halves the host->device bytes of the 8-byte record again (at 8 GPUs per host
 * the event records are what bounds the end-to-end rate) and is exact for streams on a 1 us grid -- the timestamps of every
 * DVS sensor (evfly_ros / dvs_msgs carry ros::Time, filled from the sensor's microsecond counter). H <= 511, W <= 1023,
 * n_windows <= 32768. The time of a record is a running sum, which the kernel forms per chunk of
 * evfly_accumulate_chunk_events() records: d_chunk_base_us[c] is the offset (us) from the window's first edge of the last
 * record before chunk c, chunks numbered window by window (window w has ceil(n_w / chunk) of them). The host packer that
 * writes all of this is evfly_b200.events.pack_ev4_host. Frames are those of evfly_accumulate_windows_ev8 on the same stream. */
int evfly_accumulate_chunk_events(void);
int evfly_accumulate_windows_ev4(const uint32_t* d_events, int64_t n, const int64_t* d_win_offsets, const int64_t* d_win_t0,
                                 const int64_t* d_win_t1, const int32_t* d_out_slot, const uint32_t* d_chunk_base_us,
                                 int n_windows, int H, int W, int B, int32_t* d_counts, float* d_voxel, void* d_ws,
                                 int64_t ws_bytes, void* stream);

/* counts[pol][y][x] += #events, pol 0 = negative plane, 1 = positive plane; int32 [2,H,W].
 * Events with x >= W or y >= H (unsigned compare, node.cpp:31) or polarity >= 2 are ignored.
 * The caller zeroes d_counts (or keeps accumulating into it across calls, which is how the
 * ROS callbacks between two timer ticks are served).                                      */
int evfly_accumulate_counts(const evfly_event* d_events, int64_t n, int H, int W,
                            int32_t* d_counts, void* stream);

/* Same result as evfly_accumulate_counts by a different route (opt-in experiment, measured 109 us vs 93 us for
 * 10 M events -- DESIGN.md section 8): events are binned by spatial tile
 * (2-byte records) and accumulated with shared-memory integer atomics instead of one L2 reduction per event
 * (evfly_b200/csrc/accumulate_binned.cu). d_ws: evfly_accumulate_counts_binned_workspace_bytes(n,H,W) bytes whose
 * first 256 KB must be zero on entry and are left zero on exit. Exact for any event distribution.         */
int64_t evfly_accumulate_counts_binned_workspace_bytes(int64_t n, int H, int W);
int evfly_accumulate_counts_binned(const evfly_event* d_events, int64_t n, int H, int W, int32_t* d_counts,
                                   void* d_ws, void* stream);

/* float64 frame[y][x] = pos_thresh*counts[1] - neg_thresh*counts[0], evaluated with exactly
 * that expression in IEEE double (no FMA contraction) so it is bit-identical to
 * ev_utils.py:158/:141 on the same counts.                                                */
int evfly_counts_to_frame_f64(const int32_t* d_counts, int H, int W, double pos_thresh,
                              double neg_thresh, double* d_frame, void* stream);

/* u8 publish step of the two C++ nodes (timerCallback): frame_out = apply(state_in, counts).
 * WRAP    : (state + npos - nneg) mod 256                      -- exact for any event order.
 * SATURATE: state + npos - nneg when state+npos <= 255 and state-nneg >= 0 (then no clamp can
 *           have fired whatever the order); other pixels are order dependent: they are
 *           written with the clamped order-free value, their linear index is appended to
 *           d_flagged (capacity flagged_cap) and *d_n_flagged counts them. The caller then
 *           runs evfly_u8_saturate_replay() over the window's events to make them exact.
 * d_state_in may be NULL (= all 128, the value timerCallback resets to). d_n_flagged must be
 * zero on entry in SATURATE mode.                                                         */
int evfly_counts_to_u8(const int32_t* d_counts, int H, int W, int mode,
                       const uint8_t* d_state_in, uint8_t* d_frame_out,
                       int32_t* d_flagged, int32_t flagged_cap, int32_t* d_n_flagged,
                       void* stream);

/* Exact in-order replay (evfly_dv_ros/src/node.cpp:33-41) of the pixels listed in d_flagged:
 * one CTA per flagged pixel folds the clamp-add steps of all n events in stream order.
 * n_flagged is a host value (the caller reads *d_n_flagged when it fetches the frame).     */
int evfly_u8_saturate_replay(const evfly_event* d_events, int64_t n, int H, int W,
                             const uint8_t* d_state_in, const int32_t* d_flagged,
                             int32_t n_flagged, uint8_t* d_frame_out, void* stream);

/* One window [t0_ns, t1_ns) of one stream -> int32 counts [2,H,W] (may be NULL) and a fp32
 * voxel grid [B,H,W] with bilinear weights in time (build-defined, SURVEY.md F1: the
 * reference has no voxel grid):
 *     tau_i = (B-1) * (t_i - t0) / (t1 - t0),   V[b,y,x] = sum_i pol_i * max(0, 1 - |b - tau_i|)
 * Events outside the window are ignored. d_counts / d_voxel must be zero on entry.
 * `algo`: 0 = direct (RED.s32 + 2x RED.f32 into the outputs),
 *         1 = staged (one red.global.add.v4.f32 per event into d_ws, then a finalise kernel).
 * d_ws: evfly_voxel_workspace_bytes(H,W,B) bytes, zero on entry, left zero on exit (algo 1). */
int64_t evfly_voxel_workspace_bytes(int H, int W, int B);
int evfly_voxelize_window(const evfly_event* d_events, int64_t n, int H, int W, int B,
                          int64_t t0_ns, int64_t t1_ns, int32_t* d_counts, float* d_voxel,
                          void* d_ws, int algo, void* stream);

/* T windows of one stream in one pass (utils/to_events.py:400-411 rescans the stream T times):
 * window w = [edges[w], edges[w+1]) with d_edges_ns int64[T+1] ascending.
 * d_counts int32 [T,2,H,W], d_voxel fp32 [T,B,H,W] or NULL (B ignored then).
 * Outputs need NOT be zero on entry: the call zero-fills them itself, a group of windows at
 * a time, so that a group's lines are still in L2 when its events are scattered.
 * sorted_by_time != 0 promises ascending t (ROS streams are); then each group only reads its
 * own slice of the stream. With 0 every group pass reads the whole stream (still exact).   */
int evfly_accumulate_windows(const evfly_event* d_events, int64_t n, const int64_t* d_edges_ns,
                             int T, int H, int W, int B, int32_t* d_counts, float* d_voxel,
                             int sorted_by_time, int64_t* d_range_ws /* int64[T+1] */,
                             void* stream);

/* ======================================================================================
 * L2  frame normalisation
 * ====================================================================================== */

/* evfly_ros/run.py:334-336,345-350: (u8 -> f32 - 128) * 0.2, centre crop
 * rows H/2-h/2 : H/2+h/2, cols W/2-w/2 : W/2+w/2. N frames. Also accepts int32 counts
 * [N,2,H,W] (d_u8 NULL): frame = 0.2f * (npos - nneg) in fp32, which equals the fp32 cast of
 * the reference's fp64 frame for |npos-nneg| < 2^22.                                        */
int evfly_decode_crop(const uint8_t* d_u8, const int32_t* d_counts, int N, int H, int W,
                      int h, int w, float scale, float* d_out, void* stream);

/* Per-frame q = quantile(|x|, qfrac) with torch.quantile's linear interpolation, computed
 * EXACTLY by radix-select on the fp32 bit patterns (no sort), then
 * out = clip(x / q, lo, hi) (run.py:250-253; dataloading.py:518-521) and, if cutoff > 0,
 * |out| < cutoff -> 0 (dataloading.py:531-533 / learner_models.py:477).
 * d_q (may be NULL) receives the N quantiles. One CTA per frame; the frame is re-read from L2.
 * In-place (d_out == d_x) is allowed.                                                      */
int evfly_quantile_scale_clip(const float* d_x, int N, int64_t elems_per_frame, float qfrac,
                              float lo, float hi, float cutoff, float* d_out, float* d_q,
                              void* stream);

/* ======================================================================================
 * "next" rows: the callers / formats either side of the path
 * ====================================================================================== */

/* N2: difflog event approximation (envtest/ros/run_competition.py:603-635, SMALL_EPS = 1e-5;
 * utils/to_events.py:417-439 with inputs_are_log = 1): d = log(im+eps) - log(prev+eps) in float64;
 * all zeros if max|d| < max(pos,neg), else events = (d // pos)*pos for d > 0, (d // -neg)*(-neg) for
 * d < 0, with numpy's floor_divide semantics. d_ws8: 8 bytes of scratch.                           */
int evfly_difflog_events_f64(const double* d_im, const double* d_prev, int64_t n, double eps,
                             int inputs_are_log, double pos_thresh, double neg_thresh, double* d_events,
                             void* d_ws8, void* stream);

/* N1: in-place |x| < cutoff -> 0 (learner/dataloading.py:531-533); the percentile scaling itself is
 * evfly_quantile_scale_clip.                                                                       */
int evfly_min_cutoff_f32(float* d_x, int64_t n, float cutoff, void* stream);

/* evfly_decode_crop (counts form) followed by evfly_quantile_scale_clip, fused for frames that come from integer
 * counts: out[n] = clip(0.2f*(n+ - n-) / quantile(|.|, q), lo, hi) over the centre crop, bit-identical to the two
 * calls. The order statistics are found exactly from one integer histogram of |n+ - n-| (run.py:334-336,345-351,
 * 250-253; learner/dataloading.py:515-524). d_q (optional): the N quantiles.                              */
int evfly_counts_normalise(const int32_t* d_counts, int N, int H, int W, int h, int w, float scale, float qfrac,
                           float lo, float hi, float cutoff, float* d_out, float* d_q, void* stream);

/* ---- next row N3: rectification (utils/calibration_tools/rectify_bag.py:91-138; evfly_ros/run.py:339-340) ----
 * dst[n][i][j] = cv2.remap(src[n], mapx, mapy, INTER_CUBIC) (constant-0 border), bit for bit: coordinates rounded
 * half-even to 1/32 pixel, A = -0.75 cubic table, OpenCV's summation order, no FMA. src: float32 [N,H,W], or with
 * src_is_u8 the accumulator's byte image decoded on the fly as (v - 128) * 0.2 (run.py:334-336). The maps
 * [OH rows of OW, row pitch map_ld floats] are shared by the N images; passing a window of the maps computes only
 * that window of the output (e.g. the 260x346 centre crop of run.py:346-351). flip: img[:, ::-1] before the remap;
 * rotate: cv2.rotate(ROTATE_180) of the result (remap_img's arguments).                                       */
int evfly_remap_bicubic_f32(const void* d_src, int src_is_u8, int N, int H, int W, const float* d_mapx,
                            const float* d_mapy, int64_t map_ld, int OH, int OW, int flip, int rotate, float* d_dst,
                            void* stream);

/* remap_events (rectify_bag.py:101-116): out_x = mapx[y][x], out_y = mapy[y][x] (maps [H,W]), optional rotation
 * about (target_w-1, target_h-1), keep[i] = 1 when the result lies inside [0, target-1]; the caller compacts. */
int evfly_remap_events_f32(const int32_t* d_x, const int32_t* d_y, int64_t n, const float* d_mapx,
                           const float* d_mapy, int H, int W, int rotate, int target_w, int target_h,
                           float* d_out_x, float* d_out_y, uint8_t* d_keep, void* stream);

/* ======================================================================================
 * L3  model forward: operator entry points (fp32 exact path)
 *
 * The reference's forward is stock PyTorch (learner/learner_models.py:521-616,
 * learner/vitfly_models.py:132-150, learner/ViTsubmodules.py:61-148,
 * learner/ConvLSTM_pytorch/convlstm.py:38-53); there is no native interface to mirror, so the
 * ABI exposes the operators those forwards are made of. The host-side drop-in nn.Modules
 * (evfly_b200/learner_models.py ...) keep the reference's constructors, state_dict keys and
 * forward() signatures and sequence these calls; they are capturable in a CUDA graph.
 * Tensors are fp32; layouts are described by explicit element strides so that the
 * reference's flatten/transpose/permute/cat glue costs no copies.
 * ====================================================================================== */

#define EVFLY_ACT_NONE     0
#define EVFLY_ACT_RELU     1
#define EVFLY_ACT_LEAKY    2   /* negative slope 0.01 (F.leaky_relu default)            */
#define EVFLY_ACT_GELU     3   /* exact erf GELU (nn.GELU default)                       */
#define EVFLY_ACT_TANH     4
#define EVFLY_ACT_SIGMOID  5

/* y = post_scale[c] * act(conv(x, w) + bias[c]) + post_shift[c] + res
 * Direct convolution as an implicit GEMM (M = N*OH*OW pixels, N = Cout/groups, K = Cin/groups*KH*KW).
 * x is addressed as x[n*xs[0] + c*xs[1] + h*xs[2] + w*xs[3]], y and res likewise with ys: a
 * Linear layer over tokens [rows, K] is the 1x1 case with xs = {0, 1, 0, K}.
 * w is PyTorch's [Cout, Cin/groups, KH, KW] contiguous. bias / post_* / res may be NULL.     */
typedef struct evfly_conv2d_args {
    const float* x;
    const float* w;
    const float* bias;
    const float* post_scale;
    const float* post_shift;
    const float* res;
    float*       y;
    int32_t N, Cin, H, W, Cout, KH, KW, stride, pad, groups, act, reserved;
    int64_t xs[4];
    int64_t ys[4];
} evfly_conv2d_args;
int evfly_conv2d_f32(const evfly_conv2d_args* args, void* stream);

/* y[m,n] = act(sum_k x[m,k] w[n,k] + bias[n]) (+ res[m,n]) for M <= 8 rows (batch-1 streaming):
 * one warp per output feature streams its weight row once. Rows of x / res / y may be strided.  */
int evfly_linear_smallm_f32(const float* d_x, int64_t x_ld, const float* d_w, const float* d_bias,
                            const float* d_res, int64_t res_ld, float* d_y, int64_t y_ld, int M, int N,
                            int K, int act, void* stream);

/* 2-D pooling without padding, floor mode, on contiguous [planes,H,W] -> [planes,OH,OW].
 * mode 0 = max, 1 = average. negate_in computes pool(-x) (the reference's min-pool idiom,
 * vitfly_models.py:58 / learner_models.py:76-93), negate_out negates the result.            */
int evfly_pool2d_f32(const float* d_x, float* d_y, int64_t planes, int H, int W, int k, int stride,
                     int mode, int negate_in, int negate_out, void* stream);

/* Bilinear resize (F.interpolate / nn.Upsample) of x[n*xs[0]+c*xs[1]+h*xs[2]+w*xs[3]]
 * ([N,C,H,W]) to OHxOW, written to y[n*ys[0]+c*ys[1]+oh*ys[2]+ow*ys[3]] (so it can land inside
 * a concat buffer), then y = clip(y*mul + add, lo, hi) (mul=1, add=0, lo=-inf, hi=+inf for a
 * plain resize).                                                                             */
int evfly_resize_bilinear_f32(const float* d_x, const int64_t* xs, float* d_y, const int64_t* ys,
                              int N, int C, int H, int W, int OH, int OW, int align_corners,
                              float mul, float add, float lo, float hi, void* stream);

/* resize_bilinear(clip(x * pre_mul, pre_lo, pre_hi)) without materialising the clipped tensor: the glue between
 * the depth UNet and the ViT (learner_models.py:634 `x_depth * 2` clamped to [0,1], then vitfly_models.py:28-29
 * resize to 60x90). Same values as the two-step form (the map is applied to the four samples).              */
int evfly_resize_bilinear_premap_f32(const float* d_x, const int64_t* xs, float* d_y, const int64_t* ys, int N,
                                     int C, int H, int W, int OH, int OW, int align_corners, float pre_mul,
                                     float pre_lo, float pre_hi, void* stream);

/* nn.LayerNorm over the last dimension of contiguous [rows, C] (C <= 1024), eps inside sqrt. */
int evfly_layernorm_f32(const float* d_x, const float* d_gamma, const float* d_beta, float* d_y,
                        int64_t rows, int C, float eps, void* stream);

/* softmax(q k^T / sqrt(C/heads)) v with a handful of keys (spatial-reduction attention,
 * ViTsubmodules.py:74-80): q [B,N,C], kv [B,n_kv,2C] laid out [key|value][head][d], out [B,N,C].
 * n_kv <= 32.                                                                               */
int evfly_attention_small_f32(const float* d_q, const float* d_kv, float* d_out, int B, int N,
                              int C, int heads, int n_kv, void* stream);

/* Strided 4-D elementwise map: y[idx.ys] = clip(x[idx.xs]*mul/div + add, lo, hi) over
 * dims[4]. Serves the crop-skip copy (learner_models.py:512), clip(depth*2,0,1) (:634), the
 * metadata concat desvel/10, desvel*0.1, quat (vitfly_models.py:144,65) and plain copies.
 * `div` keeps x/10 and x*0.1 distinct in fp32, as the reference has both.                    */
int evfly_map4d_f32(const float* d_x, const int64_t* xs, float* d_y, const int64_t* ys,
                    const int64_t* dims, float mul, float div, float add, float lo, float hi,
                    void* stream);

/* nn.PixelShuffle(r) of x ([N, C*r*r, H, W], strides xs) into y ([N,C,H*r,W*r], strides ys).   */
int evfly_pixel_shuffle_f32(const float* d_x, const int64_t* xs, float* d_y, const int64_t* ys,
                            int N, int C, int H, int W, int r, void* stream);

/* OrigUNet.form_input (learner_models.py:476-494): IN PLACE x[|x| < cutoff] = 0 on the caller's
 * frames [N,1,H,W] (the reference mutates its input too), then
 *   form_bev 0: out [N,2,H,W], BOTH channels = max(x,0) (the reference's expand() aliasing),
 *   form_bev 1: out = |x|,   form_bev 2: out = (x != 0) ? 1 : 0  (NaN != 0 -> 1).   */
int evfly_form_input_f32(float* d_x, float* d_out, int64_t n, int64_t plane, int form_bev,
                         float cutoff, void* stream);

/* nn.LSTM over an UNBATCHED sequence (the reference feeds [N,feat], so N is time): one layer.
 * d_gx [T,4H] = x_t W_ih^T + b_ih + b_hh (a conv2d_f32 call), gate order i,f,g,o;
 * d_whh_t [H,4H] = W_hh transposed; d_h0/d_c0 [H] (NULL = zeros); d_hs [T,H] receives every h_t;
 * d_hT/d_cT [H] the final state. One persistent CTA walks the T steps with h,c in shared memory.
 * n_seq > 1 runs that many independent sequences (trajectories of one rank), one CTA each, on
 * time-major batches: d_gx [T, n_seq, 4H], d_hs [T, n_seq, H], states [n_seq, H].               */
int evfly_lstm_seq_f32(const float* d_gx, const float* d_whh_t, const float* d_h0, const float* d_c0,
                       float* d_hs, float* d_hT, float* d_cT, int T, int H, int n_seq, void* stream);

/* LSTM cell update for n rows (gate order i,f,g,o): gates [n,4H] = x W_ih^T + h W_hh^T + biases (two
 * evfly_linear_smallm_f32 calls), c_in [n,H] or NULL (zeros) -> c_out, h_out (and h_out2 if not NULL).
 * The short-sequence / batch-1 form of evfly_lstm_seq_f32: the gate GEMV spreads over many CTAs.   */
int evfly_lstm_pointwise_f32(const float* d_gates, const float* d_c_in, float* d_c_out, float* d_h_out,
                             float* d_h_out2, int n, int H, void* stream);

/* ConvLSTM cell pointwise (convlstm.py:44-53), gate order i,f,o,g: gates [4*Ch, P],
 * c [Ch,P] updated in place, h_out [Ch,P]:  c = sig(f)*c + sig(i)*tanh(g);  h = sig(o)*tanh(c).  */
int evfly_convlstm_pointwise_f32(const float* d_gates, float* d_c, float* d_h_out, int Ch, int P,
                                 void* stream);

/* VelPredictor num_out=1 tail (learner_models.py:321-333): y [N,1] -> out [N,3] =
 * [sqrt(clip(1-y^2,0,1)), y, 0].                                                             */
int evfly_velpred_unit_f32(const float* d_y, float* d_out, int N, void* stream);

/* ======================================================================================
 * L3  model forward: bf16 tensor-core path (tcgen05 / TMEM / TMA)
 *
 * Activations are bf16 NHWC on a pitch-preserving grid [N, Hp, Wp, C]: every tensor of a UNet
 * level keeps the level's input Hp x Wp as its pitch and a valid 3x3 conv only shrinks the VALID
 * extent (vh x vw); positions outside it hold don't-care values that no valid output ever reads.
 * On that grid a 3x3 valid conv is out[m] = sum_taps in[m + kh*Wp + kw] W[tap]: nine shifted GEMMs
 * accumulated in TMEM (see evfly_b200/csrc/tc_conv_bf16.cu).
 * ====================================================================================== */

/* out[dst(m), c0 + n] = relu?( sum_{tap,k} x[m + shift(tap), k] * w[n, tap*Cin + k] + bias[n] + res[m, n] )
 *   x   bf16 [M_rows, Cin]            (the flattened pitch grid)
 *   w   bf16 [n_rows, taps*Cin]       ([Cout][tap][Cin] for a conv; [4][Cout][Cin] for convt)
 *   taps 9: shift = kh*w_pitch + kw (3x3 valid conv);  taps 1: plain GEMM (1x1 conv / Linear)
 *   out bf16 (or out_f32) with out_ld elements per destination pixel, channel offset out_c0
 *   res_f32: optional fp32 [M_rows, n_rows] added before the activation (ConvLSTM x-gates)
 *   convt = 1: ConvTranspose2d(k=2,s=2): GEMM column n = (2a+b)*cout_t + co goes to destination
 *              pixel (img, 2*ih+a, 2*iw+b) of the compact grid [N, 2*valid_h, 2*valid_w]; rows of
 *              the source grid [N,Hp,Wp] outside valid_h x valid_w are skipped.
 * Cin must be a multiple of 32.                                                              */
/* flags: EVFLY_TC_COMPACT -- write the result on a COMPACT grid [N, valid_h, valid_w, ...] instead of the source's pitch grid:
 * row (n, ih, iw) of [N,Hp,Wp] with ih < valid_h, iw < valid_w (give the OUTPUT's valid extent) goes to pixel
 * (n*valid_h + ih)*valid_w + iw, other rows are skipped. The second 3x3 conv of a UNet level then runs on (Hp-2)(Wp-2) rows
 * instead of Hp*Wp, the ConvLSTM on 8x13 instead of 12x17 positions per frame.                                              */
#define EVFLY_TC_COMPACT 1
typedef struct evfly_tc_conv_args {
    const void*  x;
    const void*  w;
    const float* bias;
    const float* res_f32;
    void*        out;
    float*       out_f32;
    int64_t M_rows;
    int64_t out_ld;
    int32_t Cin, n_rows, taps, w_pitch, relu, out_c0;
    int32_t convt, Hp, Wp, valid_h, valid_w, cout_t;
    const void*  res_bf16;   /* optional bf16 [M_rows, n_rows] residual (x + attn(x), x + ffn(x)) */
    int64_t flags;           /* 0, or EVFLY_TC_COMPACT */
    /* fused ConvLSTM cell (convlstm.py:44-53): with weight rows interleaved as n = 4*ch + gate
     * (gate order i,f,o,g) the epilogue computes c = sig(f)*c + sig(i)*tanh(g), h = sig(o)*tanh(c)
     * in place on lstm_c fp32 [M_rows, n_rows/4] and writes h as bf16 [M_rows, n_rows/4]; the gate
     * pre-activations (GEMM + res_f32 x-gates) never reach memory. out / out_f32 are ignored.      */
    float*       lstm_c;
    void*        lstm_h;
} evfly_tc_conv_args;
int evfly_tc_conv_bf16(const evfly_tc_conv_args* args, void* stream);

/* ConvLSTM (1x1 kernel, no bias) recurrence over T steps in one call (convlstm.py:136-176 loops in Python):
 * d_h_all bf16 [(T+1)*P, Ch] holds h_0 in block 0 and receives h_t in block t+1; d_gx fp32 [T*P, 4*Ch] are the
 * x-gate pre-activations, d_wh bf16 [4*Ch, Ch] the h half of the gate conv, both with gate-interleaved
 * rows/columns n = 4*ch + gate (i,f,o,g); d_c fp32 [P, Ch] is updated in place.
 * d_sync != NULL (8 bytes of device scratch): ONE persistent launch walks the T steps, the <= 148 co-resident
 * CTAs meeting at a grid-wide barrier between steps; d_sync == NULL: T fused step kernels back to back.     */
int evfly_convlstm_scan_bf16(void* d_h_all, const void* d_wh, const float* d_gx, float* d_c, int T, int64_t P,
                             int Ch, void* d_sync, void* stream);

/* The same recurrence with the x half of the gate convolution inside the step (evfly_b200/csrc/convlstm_scan.cu): d_x bf16
 * [T*P, Cx] are the ConvLSTM inputs in time-major order (learner_models.py:544-546 feeds unet_e52's output), d_wx bf16
 * [4*Ch, Cx] and d_wh bf16 [4*Ch, Ch] the two halves of the gate conv (rows n = 4*ch + gate, i,f,o,g). The x-gates never exist
 * in memory: per step gates = x_t W_x^T + h_{t-1} W_h^T is accumulated in TMEM (K = Cx + Ch) and the x part of step t+1 runs
 * while the grid waits for h_t. Always one persistent cooperative launch (d_sync: 8 bytes of device scratch); returns
 * EVFLY_ERR_UNSUPPORTED when the device cannot keep the grid co-resident -- use the x-gate GEMM + evfly_convlstm_scan_bf16
 * then. Results differ from that pair only by fp32 summation order (the x-gates are no longer rounded to fp32 in between). */
int evfly_convlstm_scan_fused_bf16(const void* d_x, const void* d_wx, void* d_h_all, const void* d_wh, float* d_c, int T,
                                   int64_t P, int Cx, int Ch, void* d_sync, void* stream);

/* First UNet layer (learner_models.py OrigUNet unet_e11, Cin = 1 or 2): fp32 NCHW [N,Cin,H,W] -> 3x3 valid
 * conv + bias + ReLU -> bf16 NHWC [N,H,W,32] on the input's own grid (valid (H-2)x(W-2)).
 * w fp32 [32,Cin,3,3] (PyTorch layout), bias fp32 [32]. Runs on tcgen05: every thread builds the im2col row
 * of its pixel (K = 9*Cin zero-padded to 32, bf16) in swizzled shared memory, two MMAs per 128-pixel tile
 * (evfly_b200/csrc/tc_stem.cu); bias is added in fp32.                                                     */
int evfly_stem_conv3x3_bf16(const float* d_x, const float* d_w, const float* d_bias, void* d_out,
                            int N, int Cin, int H, int W, void* stream);

/* The same layer on CUDA cores with fp32 inputs and weights (FMA-bound; kept as the cross-check of the tensor-core
 * stem and for inputs whose bf16 rounding matters).                                                       */
int evfly_stem_conv3x3_fma_bf16(const float* d_x, const float* d_w, const float* d_bias, void* d_out,
                                int N, int Cin, int H, int W, void* stream);

/* 2x2 stride-2 max-pool of the valid region of a bf16 NHWC pitch grid [N,Hp,Wp,C] (valid vh x vw)
 * into a compact grid [N, vh/2, vw/2, C]. C % 8 == 0.                                         */
int evfly_maxpool2x2_nhwc_bf16(const void* d_x, void* d_y, int N, int Hp, int Wp, int vh, int vw, int C,
                               void* stream);

/* Bilinear resize (align_corners=False) of the valid region of a bf16 NHWC pitch grid to OHxOW,
 * written as bf16 to y[(n*OH+oh)*OW+ow][c0 + c] with out_ld elements per pixel (the `interp` skip
 * connection landing in its half of the decoder's concat buffer). C % 8 == 0.                 */
int evfly_resize_bilinear_nhwc_bf16(const void* d_x, void* d_y, int N, int Hp, int Wp, int vh, int vw,
                                    int C, int OH, int OW, int64_t out_ld, int out_c0, void* stream);

/* Copy (crop) of a window of the valid region: y[(n*OH+oh)*OW+ow][c0+c] = x[n, h0+oh, w0+ow, c].  */
int evfly_crop_nhwc_bf16(const void* d_x, void* d_y, int N, int Hp, int Wp, int C, int h0, int w0, int OH,
                         int OW, int64_t out_ld, int out_c0, void* stream);

/* ConvLSTM cell update on the pitch grid: gates fp32 [P, 4*Ch] pixel-major, gate order i,f,o,g;
 * c fp32 [P,Ch] in place; h written as bf16 [P,Ch] (next step's GEMM operand + decoder input).  */
int evfly_convlstm_pointwise_nhwc(const float* d_gates, float* d_c, void* d_h_bf16, int64_t P, int Ch,
                                  void* stream);

/* fp32 <-> bf16 layout converters between the reference's NCHW tensors and the pitch grids:
 * to_nhwc: x fp32 NCHW [N,C,vh,vw] -> bf16 [N,Hp,Wp,C] (positions outside vh x vw zeroed);
 * to_nchw: bf16 or fp32 (src_is_f32) pitch grid -> fp32 NCHW [N,C,vh,vw].                      */
int evfly_nchw_f32_to_nhwc_bf16(const float* d_x, void* d_y, int N, int C, int vh, int vw, int Hp, int Wp,
                                void* stream);
int evfly_nhwc_to_nchw_f32(const void* d_x, int src_is_f32, float* d_y, int N, int C, int vh, int vw, int Hp,
                           int Wp, void* stream);

/* ---- bf16 path of the ViT / ViT-LSTM stages (tokens bf16 [B, N, C] = NHWC on the H' x W' grid) ---- */

/* OverlapPatchMerging / the attention's reduction conv (ViTsubmodules.py:15-34, 43-44, 68-70):
 * conv(k, stride, pad) + bias + LayerNorm(Cout) fused, one warp per output token.
 * x: fp32 NCHW with Cin == 1 (x_is_f32_nchw) or bf16 NHWC [B,H,W,Cin]; w_kc fp32 [k*k*Cin][Cout]
 * with K index (kh*k + kw)*Cin + ci; tokens bf16 [B, OH*OW, Cout]; Cout in {32, 64}.
 * From 1024 tokens up the two OverlapPatchMerging shapes (7x7/4 on the fp32 image, 3x3/2 on 32 channels) run on
 * tcgen05: im2col rows built in swizzled shared memory by the token's thread, LayerNorm in the TMEM epilogue
 * (evfly_b200/csrc/tc_patch_embed.cu); image and weights are rounded to bf16 there.                 */
int evfly_patch_embed_ln_bf16(const void* d_x, int x_is_f32_nchw, const float* d_w_kc, const float* d_bias,
                              const float* d_gamma, const float* d_beta, void* d_tokens, int B, int H, int W,
                              int Cin, int Cout, int k, int stride, int pad, float eps, void* stream);
int evfly_layernorm_bf16(const void* d_x, const float* d_gamma, const float* d_beta, void* d_y, int64_t rows,
                         int C, float eps, void* stream);
/* q bf16 [B,N,C], kv bf16 [B,n_kv,2C] ([key|value][head][d]) -> out bf16 [B,N,C]; n_kv <= 8.      */
int evfly_attention_small_bf16(const void* d_q, const void* d_kv, void* d_out, int64_t B, int N, int C,
                               int heads, int n_kv, void* stream);
/* MixFFN's grouped conv (groups = C, 8 -> 8 per group, 3x3, 'same') + bias + exact GELU on bf16
 * NHWC [B,H,W,Ce]; w is PyTorch's fp32 [Ce, 8, 3, 3] (ViTsubmodules.py:92,113).                   */
int evfly_dwconv3x3_gelu_nhwc_bf16(const void* d_x, const float* d_w, const float* d_bias, void* d_y, int64_t B,
                                   int H, int W, int Ce, void* stream);
/* evfly_lstm_seq_f32 with W_hh resident in shared memory as bf16: d_whh_pairs is bf16x2
 * [H/2][4H] = {W_hh[r][2j], W_hh[r][2j+1]} at [j][r]; state and gates stay fp32. 4H <= 1024.       */
int evfly_lstm_seq_smemw(const float* d_gx, const void* d_whh_pairs, const float* d_h0, const float* d_c0,
                         float* d_hs, float* d_hT, float* d_cT, int T, int H, int n_seq, void* stream);

/* 3x3 valid conv + bias (+ReLU) for Cin, Cout in {32, 64} (and 64 -> 128; Cin = 128 -> 64|128|256 with the weights
 * streamed through a ring instead of resident) with the input halo reused from shared memory
 * (one 4-D TMA box [18 x 10 pixels] per 16x8 output tile, all 9 taps read through shifted UMMA
 * descriptors, weights resident): x bf16 [N,Hp,Wp,Cin] valid vh x vw, w bf16 [Cout, 9*Cin] ([Cout][tap][Cin]),
 * out bf16 [N,Hp,Wp,Cout]; only the valid (vh-2) x (vw-2) outputs are written.                         */
int evfly_tc_conv3x3_halo_bf16(const void* d_x, const void* d_w, const float* d_bias, void* d_out, int N,
                               int Hp, int Wp, int vh, int vw, int Cin, int Cout, int relu, void* stream);
/* The same conv writing a COMPACT grid: d_out bf16 [N, vh-2, vw-2, Cout] (pitch = valid extent, cf. EVFLY_TC_COMPACT): for layers
 * whose consumer is the generic kernel or a transposed conv, which compute every row of the grid they are given.            */
int evfly_tc_conv3x3_halo_compact_bf16(const void* d_x, const void* d_w, const float* d_bias, void* d_out, int N,
                                       int Hp, int Wp, int vh, int vw, int Cin, int Cout, int relu, void* stream);

/* The same conv with nn.MaxPool2d(2) (learner_models.py OrigUNet pool1..pool4) fused into the epilogue: besides
 * d_out it writes d_pool bf16 [N,Hp2,Wp2,Cout], valid ((vh-2)/2) x ((vw-2)/2), bit-identical to
 * evfly_maxpool2x2_nhwc_bf16 applied to d_out (saves re-reading the full-resolution activation).       */
int evfly_tc_conv3x3_halo_pool_bf16(const void* d_x, const void* d_w, const float* d_bias, void* d_out, void* d_pool,
                                    int N, int Hp, int Wp, int vh, int vw, int Cin, int Cout, int relu, int Hp2, int Wp2,
                                    void* stream);

/* 3x3 conv with padding=1 (+bias, +ReLU) on a DENSE bf16 NHWC tensor [N,H,W,Cin] -> [N,H,W,Cout] through the same
 * halo-reuse kernel: the halo box starts at (-1,-1) and TMA zero-fills everything outside the image
 * (vitfly_models.py:120 down_sample = Conv2d(48, 12, 3, padding=1), channels zero-padded to 64 / 32).   */
int evfly_tc_conv3x3_same_bf16(const void* d_x, const void* d_w, const float* d_bias, void* d_out, int N, int H, int W,
                               int Cin, int Cout, int relu, void* stream);

/* The same conv followed by a 1x1 conv to ONE channel in the epilogue (unet_d42 + unet_out, learner_models.py:581-583):
 * out1[n,oh,ow] = b1 + sum_c bf16(relu?(conv)[c]) * w1[c], fp32 on the [N,Hp,Wp] pitch grid; the Cout-channel activation
 * is never written. d_w1 fp32 [Cout] (the 1x1 weights rounded to bf16, stored as fp32), d_b1 fp32 [1]. Cin, Cout in {32, 64}. */
int evfly_tc_conv3x3_halo_out1_bf16(const void* d_x, const void* d_w, const float* d_bias, const float* d_w1,
                                    const float* d_b1, float* d_out1, int N, int Hp, int Wp, int vh, int vw, int Cin,
                                    int Cout, int relu, void* stream);

/* vitfly_models.py:136-142: cat([PixelShuffle(2)(s2), Upsample((2*H2,2*W2), bilinear, align_corners=True)(s1)], 1)
 * from the bf16 token tensors t2 [B,H2,W2,C2], t1 [B,H1,W1,C1] into one dense bf16 NHWC tensor
 * [B,2*H2,2*W2,ld]: channels [0,C2/4) shuffle, [C2/4,C2/4+C1) upsample, [.., ld) zero.                 */
int evfly_shuffle_upsample_cat_bf16(const void* d_t2, int H2, int W2, int C2, const void* d_t1, int H1, int W1,
                                    int C1, void* d_out, int64_t B, int ld, void* stream);

/* One CTA per sample: a whole MixFFN block (learner/ViTsubmodules.py:85-120) + the residual add and the LayerNorm of
 * MixTransformerEncoderLayer.forward (:143-146) on the tensor cores with the 8C-wide activation resident on chip:
 *   out = LayerNorm(x + mlp2(GELU(groupedconv3x3(mlp1(x)))))
 * d_tokens / d_out bf16 [B, H*W, C]; d_w_img = the pre-swizzled operand images of the block, evfly_vit_ffn_image_bytes(C)
 * bytes (evfly_b200/tc.py::pack_vit_ffn); d_fbias fp32 = mlp1 bias [8C], conv bias [8C], mlp2 bias [C], LayerNorm
 * gamma [C], beta [C]. Instantiated for the two stages of LSTMNetVIT / ViT: (H,W,C) = (15,23,32) and (8,12,64).        */
int64_t evfly_vit_ffn_image_bytes(int C);
int evfly_vit_ffn_bf16(const void* d_tokens, const void* d_w_img, const float* d_fbias, void* d_out, int B, int H, int W,
                       int C, float eps, void* stream);

/* EfficientSelfAttention after the spatial reduction (ViTsubmodules.py:74-83) + the residual of :144 in one launch:
 *   out = x + finalLayer(softmax(query(x) K^T / sqrt(d)) V)
 * d_x / d_out bf16 [B, N, C]; d_kv bf16 [B, n_kv, 2C] = keyValueExtractor(LayerNorm(cn1(x))), layout [kv][head][d];
 * d_w_img: pre-swizzled images of query.weight then finalLayer.weight (2*C*C*2 bytes, tc.pack_vit_attn);
 * d_bias fp32: query bias [C], finalLayer bias [C]. (C, heads) = (32, 1) or (64, 2); n_kv <= 8.                 */
int evfly_vit_attn_bf16(const void* d_x, const void* d_kv, const void* d_w_img, const float* d_bias, void* d_out, int64_t B,
                        int N, int C, int heads, int n_kv, void* stream);

/* Level 1 of OrigUNet for a BINARY input (form_BEV = 2, learner_models.py:489-491; every shipped configuration):
 * unet_e11 (1 -> 32, 3x3 valid, bias, ReLU) has 2^9 possible outputs per pixel, so it runs as a table lookup inside
 * the unet_e12 kernel and its output never exists in HBM (evfly_b200/csrc/tc_conv_halo.cu).
 *   evfly_stem_patterns          mask fp32 [N,1,H,W] (0 / non-0) -> uint16 [N,H-2,W-2]: the 9 mask bits under each e11 pixel
 *   evfly_tc_stem_e12_pool_bf16  patterns + unet_e11.{weight [32,1,3,3], bias} + packed unet_e12 weights / bias
 *       -> y_e1 bf16 NHWC on the [N,H,W,32] pitch grid (valid (H-4) x (W-4)) and, when d_pool is given, MaxPool2d(2) of it
 *       on the [N,Hp2,Wp2,32] grid (learner_models.py:533-535).                                                      */
int evfly_stem_patterns(const float* d_mask, uint16_t* d_pat, int N, int H, int W, void* stream);
/* form_input(form_BEV = 2) + evfly_stem_patterns in one pass over the normalised frames fp32 [N,1,H,W]: |x| < cutoff is zeroed
 * IN PLACE like the reference does (learner_models.py:477) and the patterns are those of the mask x != 0 (NaN counts as set). */
int evfly_form_patterns(float* d_frames, float cutoff, uint16_t* d_pat, int N, int H, int W, void* stream);
int evfly_tc_stem_e12_pool_bf16(const uint16_t* d_pat, const float* d_stem_w, const float* d_stem_b, const void* d_w,
                                const float* d_bias, void* d_out, void* d_pool, int N, int H, int W, int relu, int Hp2,
                                int Wp2, void* stream);
/* The two fused-pool convs when d_out (y_e1 / y_e2 / y_e3) is only going to be sampled by evfly_resize_bilinear_nhwc_bf16 to a
 * height of skip_OH -- the decoder's 'interp' skip (learner_models.py:512-519), its only other consumer being the pooled
 * output written here: output rows that resize does not read are NOT written (their content in d_out is undefined). With the
 * deployed sizes that is 44-52 % of the full-resolution stores. d_pool and the written rows are those of the plain calls. */
int evfly_tc_stem_e12_pool_rows_bf16(const uint16_t* d_pat, const float* d_stem_w, const float* d_stem_b, const void* d_w,
                                     const float* d_bias, void* d_out, void* d_pool, int N, int H, int W, int relu, int Hp2,
                                     int Wp2, int skip_OH, void* stream);
int evfly_tc_conv3x3_halo_pool_rows_bf16(const void* d_x, const void* d_w, const float* d_bias, void* d_out, void* d_pool,
                                         int N, int Hp, int Wp, int vh, int vw, int Cin, int Cout, int relu, int Hp2,
                                         int Wp2, int skip_OH, void* stream);

/* ======================================================================================
 * STAGE-LEVEL entry points (evfly_b200/csrc/stages.cu): whole stages of the path enqueued from C++ on the caller's
 * stream -- what the C++ side of evfly_ros (or any host language) binds instead of the per-operator functions
 * above. Weights arrive packed in caller-owned device memory, intermediates live in a caller-owned workspace,
 * nothing is allocated and nothing synchronises, so a whole step is capturable in one CUDA graph.
 * ====================================================================================== */

/* evfly_ros/run.py:334-350 + 250-253: (u8 - 128) * 0.2 (or 0.2 * (n+ - n-) from int32 count frames [N,2,H,W]),
 * centre crop to h x w, per-frame 97th percentile of |x| (torch.quantile semantics), clip(x / q, -1, 1).
 * Exactly one of d_u8 [N,H,W] / d_counts is given. d_frames fp32 [N,1,h,w].                                   */
int evfly_prep_frame(const uint8_t* d_u8, const int32_t* d_counts, int N, int H, int W, int h, int w,
                     float* d_frames, void* stream);

/* Packed weights of OrigUNet in the shipped configuration (learner/configs/*.txt:39-47: form_BEV = 2,
 * skip_type = interp, num_recurrent = [1, 0]); layouts as produced by evfly_b200/tc.py:
 *   e11_w fp32 [32,1,3,3], e11_b fp32 [32]                                            (unet_e11, table lookup)
 *   conv_w[i] bf16 [Cout][9*Cin] (K index = (kh*3+kw)*Cin + ci), conv_b[i] fp32 [Cout], i = e12, e21, e22, e31,
 *       e32, e41, e42, e51, e52, d11, d12, d21, d22, d31, d32, d41, d42
 *   up_w[l] bf16 [4*Cout][Cin] (row = (2a+b)*Cout + co of ConvTranspose2d weight [Cin,Cout,a,b]), up_b[l] fp32 [Cout]
 *   out_w fp32 [32] (unet_out.weight rounded to bf16, stored as fp32), out_b fp32 [1]      (unet_out, in d42's epilogue)
 *   lstm_wx / lstm_wh bf16 [2048][512]: the x / h halves of lstm.cell_list.0.conv.weight with rows interleaved
 *       n = 4*ch + gate (gate order i,f,o,g, convlstm.py:44)                                                  */
typedef struct evfly_unet_weights {
    const float* e11_w;
    const float* e11_b;
    const void*  conv_w[17];
    const float* conv_b[17];
    const void*  up_w[4];
    const float* up_b[4];
    const float* out_w;
    const float* out_b;
    const void*  lstm_wx;
    const void*  lstm_wh;
} evfly_unet_weights;

/* OrigUNet.forward (learner/learner_models.py:521-585) for N frames = n_traj trajectories x T steps in TIME-MAJOR
 * order (frame t*n_traj + s; n_traj = 1: the reference's one sequence): form_input (in place on d_frames, like the
 * reference) -> encoder -> ConvLSTM over T (state in: d_h0 / d_c0 fp32 [n_traj,512,vh,vw] or both NULL; state out:
 * d_hT / d_cT, each may be NULL) -> decoder -> d_y_upconv fp32 [N,1,.,.] and d_depth fp32 [N,1,H,W] (y_interp).
 * bf16 tensor-core path. d_ws: evfly_unet_workspace_bytes(N, n_traj, H, W) bytes of scratch.                   */
int64_t evfly_unet_workspace_bytes(int N, int n_traj, int H, int W);
int evfly_unet_forward(const evfly_unet_weights* weights, float* d_frames, int N, int n_traj, int H, int W, float cutoff,
                       const float* d_h0, const float* d_c0, float* d_hT, float* d_cT, float* d_depth,
                       float* d_y_upconv, void* d_ws, int64_t ws_bytes, void* stream);

/* Packed weights of LSTMNetVIT (learner/vitfly_models.py:111-150), layouts of evfly_b200/tc.py and _modbase.py:
 *   per stage: patch_w fp32 [k*k*Cin][C] (K index (kh*k+kw)*Cin + ci), patch_b, LayerNorm gamma / beta fp32 [C]
 *   per layer: red_w fp32 [r*r*C][C], red_b, ln1 gamma / beta; kv_w bf16 [2C][C], kv_b fp32 [2C];
 *              attn_img / attn_bias (tc.pack_vit_attn), ffn_img / ffn_bias (tc.pack_vit_ffn)
 *   ds_w bf16 [32][9*64] / ds_b fp32 [32] (down_sample 48 -> 12 zero-padded), dec_w bf16 [512][16*24*32] (decoder Linear with
 *   spectral norm folded, re-indexed to NHWC), dec_b fp32 [512]
 *   LSTM layer l: w_ih fp32 [512][in], b fp32 [512] (b_ih + b_hh), whh_pairs bf16x2 [64][512], whh_t fp32 [128][512]
 *   fc2_w fp32 [3][128] (spectral norm folded), fc2_b fp32 [3]                                                     */
typedef struct evfly_vit_layer_weights {
    const float* red_w; const float* red_b; const float* red_ln_g; const float* red_ln_b;
    const void*  kv_w;  const float* kv_b;
    const void*  attn_img; const float* attn_bias;
    const void*  ffn_img;  const float* ffn_bias;
} evfly_vit_layer_weights;
typedef struct evfly_vit_stage_weights {
    const float* patch_w; const float* patch_b; const float* patch_ln_g; const float* patch_ln_b;
    evfly_vit_layer_weights layer[2];
} evfly_vit_stage_weights;
typedef struct evfly_vit_lstm_weights {
    evfly_vit_stage_weights stage[2];
    const void*  ds_w;  const float* ds_b;
    const void*  dec_w; const float* dec_b;
    const float* lstm_w_ih[3]; const float* lstm_b[3]; const void* lstm_whh_pairs[3]; const float* lstm_whh_t[3];
    const float* fc2_w; const float* fc2_b;
} evfly_vit_lstm_weights;

/* LSTMNetVIT.forward for N >= 8 frames = n_traj sequences x T steps, time-major: d_depth fp32 [N,1,H,W] (premap_clamp:
 * clamp(2 d, 0, 1) first -- learner_models.py:634), d_desvel fp32 [N], d_quat fp32 [N,4] or NULL ([1,0,0,0]);
 * LSTM state in d_h0 / d_c0 fp32 [3,n_traj,128] or both NULL, out d_hT / d_cT; d_vel fp32 [N,3].                      */
int64_t evfly_vit_lstm_workspace_bytes(int N);
int evfly_vit_lstm_forward(const evfly_vit_lstm_weights* weights, const float* d_depth, int N, int n_traj, int H, int W,
                           int premap_clamp, const float* d_desvel, const float* d_quat, const float* d_h0,
                           const float* d_c0, float* d_hT, float* d_cT, float* d_vel, void* d_ws, int64_t ws_bytes,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EVFLY_B200_H */
