/*
 * ev_oracle.c -- plain-C restatement of the reference's event accumulators. TEST
 * INFRASTRUCTURE: the checker for the CUDA kernels, never the product path.
 *
 * Follows (paths relative to the evfly tree):
 *   evfly_ros/src/node.cpp:29-39      per-event update of the 128-biased uint8 image, wrapping
 *   evfly_dv_ros/src/node.cpp:29-44   same with the "< 255" / "> 0" guards (saturating)
 *   utils/ev_utils.py:155-158         two histograms (pos / neg) of one event list
 *   utils/to_events.py:400-411        one pair of histograms per window [t_start, t_end)
 * and, for the voxel grid (build-defined, the reference has none -- SURVEY.md F1):
 *   tau = (B-1)(t-t0)/(t1-t0),  V[b,y,x] = sum_i pol_i * max(0, 1 - |b - tau_i|)  in double.
 *
 * The two ROS nodes themselves cannot be compiled here (roscpp / prophesee_event_msgs /
 * dv_ros_msgs headers are not in the image), so there is no oracle/_ref for them; the loops
 * below are the ROS-free restatement of their callbacks.
 *
 * Record layout = dv_ros_msgs::Event in memory (include/evfly_b200.h).
 */
#include <stdint.h>
#include <string.h>
#include <math.h>

typedef struct {
    uint16_t x, y;
    uint32_t ts_sec, ts_nsec;
    uint8_t polarity, pad[3];
} ev_t;

static int64_t ev_time(const ev_t* e) { return (int64_t)e->ts_sec * 1000000000ll + (int64_t)e->ts_nsec; }

/* eventArrayCallback of both nodes; image is the node's image_data_ (W*H uint8, starts at 128). */
void oracle_node_accumulate(const ev_t* ev, int64_t n, int W, int H, int saturate, uint8_t* image) {
    for (int64_t i = 0; i < n; ++i) {
        const ev_t* e = &ev[i];
        if (e->polarity > 1) continue; /* packer's skip marker; a sensor never emits it */
        if (e->x < W && e->y < H) {
            uint8_t* px = &image[(int64_t)e->y * W + e->x];
            if (e->polarity == 1) {
                if (!saturate) (*px)++;            /* evfly_ros: uint8 wraps 255 -> 0   */
                else if (*px < 255) (*px)++;       /* evfly_dv_ros: prevent overflow    */
            } else {
                if (!saturate) (*px)--;            /* wraps 0 -> 255                    */
                else if (*px > 0) (*px)--;         /* prevent underflow                 */
            }
        }
    }
}

/* counts[0] = negative events per pixel, counts[1] = positive events per pixel ([2,H,W]). */
void oracle_counts(const ev_t* ev, int64_t n, int H, int W, int32_t* counts) {
    for (int64_t i = 0; i < n; ++i) {
        const ev_t* e = &ev[i];
        if (e->polarity > 1 || e->x >= W || e->y >= H) continue;
        counts[((int64_t)e->polarity * H + e->y) * W + e->x] += 1;
    }
}

/* One window [t0,t1): counts [2,H,W] (may be NULL) and voxel [B,H,W] in double. */
void oracle_voxel(const ev_t* ev, int64_t n, int H, int W, int B, int64_t t0, int64_t t1,
                  int32_t* counts, double* voxel) {
    for (int64_t i = 0; i < n; ++i) {
        const ev_t* e = &ev[i];
        if (e->polarity > 1 || e->x >= W || e->y >= H) continue;
        const int64_t t = ev_time(e);
        if (t < t0 || t >= t1) continue;
        if (counts) counts[((int64_t)e->polarity * H + e->y) * W + e->x] += 1;
        if (!voxel) continue;
        const double pol = e->polarity ? 1.0 : -1.0;
        const double tau = (double)(B - 1) * (double)(t - t0) / (double)(t1 - t0);
        for (int b = 0; b < B; ++b) {
            const double w = 1.0 - fabs((double)b - tau);
            if (w > 0.0) voxel[((int64_t)b * H + e->y) * W + e->x] += pol * w;
        }
    }
}

/* T windows of one stream, each scanned over the WHOLE stream like to_events.py:405-406.
 * counts [T,2,H,W]; voxel [T,B,H,W] double or NULL. */
void oracle_windows(const ev_t* ev, int64_t n, const int64_t* edges, int T, int H, int W, int B,
                    int32_t* counts, double* voxel) {
    for (int w = 0; w < T; ++w) {
        if (edges[w + 1] <= edges[w]) continue;
        oracle_voxel(ev, n, H, W, B, edges[w], edges[w + 1], counts + (int64_t)w * 2 * H * W,
                     voxel ? voxel + (int64_t)w * B * H * W : 0);
    }
}

/* sum over all events of |w_lo| + |w_hi| per voxel cell: the scale the fp32 error bound is
 * stated against (tests use |V_gpu - V_oracle| <= 1e-6 * max(1, sum|w|)). */
void oracle_voxel_abs(const ev_t* ev, int64_t n, int H, int W, int B, int64_t t0, int64_t t1,
                      double* voxel_abs) {
    for (int64_t i = 0; i < n; ++i) {
        const ev_t* e = &ev[i];
        if (e->polarity > 1 || e->x >= W || e->y >= H) continue;
        const int64_t t = ev_time(e);
        if (t < t0 || t >= t1) continue;
        const double tau = (double)(B - 1) * (double)(t - t0) / (double)(t1 - t0);
        for (int b = 0; b < B; ++b) {
            const double w = 1.0 - fabs((double)b - tau);
            if (w > 0.0) voxel_abs[((int64_t)b * H + e->y) * W + e->x] += w;
        }
    }
}
