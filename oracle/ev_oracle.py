"""numpy / C restatement of the reference's event accumulation and frame normalisation.

TEST INFRASTRUCTURE: the checker for the CUDA kernels (tests/, smoke(), bench.py's cpu_baseline
and --impl reference legs). Never imported by evfly_b200/.

Pinned against the reference itself: tests/golden/make_golden_events.py imports
/root/reference/utils/ev_utils.py (in the authoring container) and stores its outputs on seeded
inputs in tests/golden/events_golden.npz; tests/test_oracle_events.py checks this file against
those vectors, and against numpy.histogram2d / torch.quantile, which ARE the third-party
arithmetic the reference calls.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NS = 1_000_000_000

EVENT_DTYPE = np.dtype([
    ("x", "<u2"), ("y", "<u2"), ("ts_sec", "<u4"), ("ts_nsec", "<u4"),
    ("polarity", "u1"), ("pad", "u1", (3,)),
])

_clib = None


def clib() -> C.CDLL:
    global _clib
    if _clib is None:
        from . import build as _b
        lib = C.CDLL(_b.build())
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        lib.oracle_node_accumulate.argtypes = [vp, i64, i32, i32, i32, vp]
        lib.oracle_counts.argtypes = [vp, i64, i32, i32, vp]
        lib.oracle_voxel.argtypes = [vp, i64, i32, i32, i32, i64, i64, vp, vp]
        lib.oracle_windows.argtypes = [vp, i64, vp, i32, i32, i32, i32, vp, vp]
        lib.oracle_voxel_abs.argtypes = [vp, i64, i32, i32, i32, i64, i64, vp]
        for f in (lib.oracle_node_accumulate, lib.oracle_counts, lib.oracle_voxel, lib.oracle_windows,
                  lib.oracle_voxel_abs):
            f.restype = None
        _clib = lib
    return _clib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _rec(records) -> np.ndarray:
    r = np.ascontiguousarray(records)
    assert r.dtype == EVENT_DTYPE, r.dtype
    return r


# ---------------------------------------------------------------------------------------------
# utils/ev_utils.py:113-161  form_eventframe
# ---------------------------------------------------------------------------------------------
def hist2d_counts(xs, ys, H, W) -> np.ndarray:
    """np.histogram2d(xs, ys, bins=(W,H), range=[[0,W],[0,H]])[0].T restated (ev_utils.py:141,158-159):
    unit bins, bin = floor(coord); the right edge belongs to the last bin; the rest is dropped."""
    xs = np.asarray(xs, dtype=np.float64)
    ys = np.asarray(ys, dtype=np.float64)
    ok = (xs >= 0) & (xs <= W) & (ys >= 0) & (ys <= H)  # NaN compares false
    bx = np.minimum(np.floor(xs[ok]).astype(np.int64), W - 1)
    by = np.minimum(np.floor(ys[ok]).astype(np.int64), H - 1)
    return np.bincount(by * W + bx, minlength=H * W).reshape(H, W)


def event_counts_rows(rows, H, W, neg_is_zero: bool) -> np.ndarray:
    """int64 [2,H,W] (plane 0 negative, plane 1 positive) from float rows (t,x,y,...,p)."""
    rows = np.asarray(rows)
    p = rows[:, -1]
    pos = rows[p > 0]                                     # ev_utils.py:137 / :155
    neg = rows[p == 0] if neg_is_zero else rows[p < 0]    # ev_utils.py:156 / :138
    return np.stack([hist2d_counts(neg[:, 1], neg[:, 2], H, W),
                     hist2d_counts(pos[:, 1], pos[:, 2], H, W)])


def counts_to_frame(counts, pos_thresh=0.2, neg_thresh=0.2) -> np.ndarray:
    """ev_utils.py:158: pos_thresh*hist_pos - neg_thresh*hist_neg in float64 (two products, one
    subtraction; NOT thresh*(npos-nneg), which differs by 1 ulp -- SURVEY.md F8g)."""
    return pos_thresh * counts[1].astype(np.float64) - neg_thresh * counts[0].astype(np.float64)


def form_eventframe(view_events, H, W, times0=None, times1=None, N=None, pos_thresh=0.2,
                    neg_thresh=0.2, all_events=False):
    """Restatement of utils/ev_utils.py:113-161 (same branches, same return values)."""
    view_events = np.asarray(view_events)
    if not all_events:
        if len(view_events) == 0:                                         # :118-119
            return np.zeros((H, W)), times0
        if times0 is None:                                                # :121-123
            raise SystemExit
        t = view_events[:, 0]
        if times1 is not None:                                            # :125-129
            timed = view_events[(t >= times0 * 1e9) & (t < times1[0] * 1e9)]
        elif N is not None:                                               # :130-133
            timed = view_events[t >= times0 * 1e9][:N]
            times1 = (timed[-1, 0] + 1) / 1e9
        else:                                                             # :134-135
            raise ValueError("form_eventframe() requires either times1 or N to be not None")
        counts = event_counts_rows(timed, H, W, neg_is_zero=False) if len(timed) else np.zeros((2, H, W), np.int64)
        return counts_to_frame(counts, pos_thresh, neg_thresh), times1
    if len(view_events) == 0:                                             # :152-153
        return np.zeros((H, W))
    counts = event_counts_rows(view_events, H, W, neg_is_zero=True)
    return counts_to_frame(counts, pos_thresh, neg_thresh)


# ---------------------------------------------------------------------------------------------
# packed records: counts, node accumulators, voxel grids, windows (C loops)
# ---------------------------------------------------------------------------------------------
def event_counts(records, H, W) -> np.ndarray:
    r = _rec(records)
    counts = np.zeros((2, H, W), dtype=np.int32)
    clib().oracle_counts(_p(r), r.shape[0], H, W, _p(counts))
    return counts


def node_accumulate(records, saturate: bool, W=640, H=480, image=None) -> np.ndarray:
    """evfly_ros/src/node.cpp:29-39 (wrap) / evfly_dv_ros/src/node.cpp:29-44 (saturate) applied to
    `image` (default: the value timerCallback resets to, 128). Returns the flat uint8 image."""
    r = _rec(records)
    if image is None:
        image = np.full(W * H, 128, dtype=np.uint8)
    else:
        image = np.ascontiguousarray(image, dtype=np.uint8).copy().reshape(-1)
    clib().oracle_node_accumulate(_p(r), r.shape[0], W, H, int(bool(saturate)), _p(image))
    return image


def voxel_window(records, H, W, B, t0_ns, t1_ns, want_abs=False):
    """(counts int32 [2,H,W], voxel float64 [B,H,W][, sum|w| float64 [B,H,W]])."""
    r = _rec(records)
    counts = np.zeros((2, H, W), dtype=np.int32)
    voxel = np.zeros((B, H, W), dtype=np.float64)
    clib().oracle_voxel(_p(r), r.shape[0], H, W, B, int(t0_ns), int(t1_ns), _p(counts), _p(voxel))
    if not want_abs:
        return counts, voxel
    vabs = np.zeros((B, H, W), dtype=np.float64)
    clib().oracle_voxel_abs(_p(r), r.shape[0], H, W, B, int(t0_ns), int(t1_ns), _p(vabs))
    return counts, voxel, vabs


def windows(records, edges_ns, H, W, B=None):
    """(counts int32 [T,2,H,W], voxel float64 [T,B,H,W] or None); to_events.py:400-411."""
    r = _rec(records)
    edges = np.ascontiguousarray(edges_ns, dtype=np.int64)
    T = edges.shape[0] - 1
    counts = np.zeros((T, 2, H, W), dtype=np.int32)
    voxel = np.zeros((T, B, H, W), dtype=np.float64) if B else None
    clib().oracle_windows(_p(r), r.shape[0], _p(edges), T, H, W, B or 1, _p(counts), _p(voxel))
    return counts, voxel


def sliced_frames(x, y, t, p, t_edges, H, W, pos_thresh=0.2, neg_thresh=0.2) -> np.ndarray:
    """utils/to_events.py:400-411 restated literally: per window, boolean masks over the whole
    stream, two histograms, frames[i] = frame.T."""
    x, y, t, p = (np.asarray(v) for v in (x, y, t, p))
    T = len(t_edges) - 1
    frames = np.zeros((T, H, W))
    for i in range(T):
        in_w = (t >= t_edges[i]) & (t < t_edges[i + 1])
        pos, neg = in_w & (p > 0), in_w & (p < 0)
        frames[i] = pos_thresh * hist2d_counts(x[pos], y[pos], H, W) - neg_thresh * hist2d_counts(x[neg], y[neg], H, W)
    return frames


# ---------------------------------------------------------------------------------------------
# L2: evfly_ros/run.py:334-350 decode + crop; run.py:250-253 / dataloading.py:518-533 normalise
# ---------------------------------------------------------------------------------------------
def decode_crop(u8_frame, h=260, w=346) -> np.ndarray:
    """run.py:334-336,345-350: float32 (u8 - 128) * 0.2, centre crop."""
    f = np.asarray(u8_frame).astype(np.float32)
    f -= 128
    f *= 0.2
    H, W = f.shape[-2:]
    if H != h or W != w:
        f = f[..., H // 2 - h // 2: H // 2 + h // 2, W // 2 - w // 2: W // 2 + w // 2]
    return f


def quantile_f32(absx: np.ndarray, q: float) -> np.float32:
    """torch.quantile(x, q) for a 1-D float32 tensor, 'linear' interpolation, restated from
    ATen's quantile_compute: rank = q*(n-1) in float32, lerp(sorted[floor], sorted[ceil], frac)
    with at::lerp's two-sided formula."""
    v = np.sort(np.asarray(absx, dtype=np.float32).reshape(-1))
    n = v.shape[0]
    rank = np.float32(q) * np.float32(n - 1)
    below, above = np.floor(rank), np.ceil(rank)
    wgt = np.float32(rank - below)
    a, b = v[int(below)], v[int(above)]
    diff = np.float32(b - a)
    if wgt < np.float32(0.5):
        return np.float32(a + np.float32(wgt * diff))
    return np.float32(b - np.float32(diff * np.float32(np.float32(1.0) - wgt)))


def quantile_scale_clip(frames, q=0.97, lo=-1.0, hi=1.0, cutoff=0.0):
    """Per frame: s = quantile(|x|, q); clip(x / s, lo, hi); |.| < cutoff -> 0.
    run.py:250-253 (one frame) and dataloading.py:518-521,531-533 (per frame of a trajectory).
    Returns (out float32 like frames, s float32 [N])."""
    x = np.asarray(frames, dtype=np.float32)
    flat = x.reshape(x.shape[0], -1)
    out = np.empty_like(flat)
    qs = np.empty(flat.shape[0], dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(flat.shape[0]):
            s = quantile_f32(np.abs(flat[i]), q)
            qs[i] = s
            v = flat[i] / s
            v = np.where(np.isnan(v), v, np.clip(v, np.float32(lo), np.float32(hi)))
            if cutoff > 0:
                v = np.where(np.abs(v) < np.float32(cutoff), np.float32(0), v)
            out[i] = v
    return out.reshape(x.shape), qs


# ---------------------------------------------------------------------------------------------
# "next" rows: difflog events (run_competition.py:603-635, to_events.py:417-439), dataset normalisation
# ---------------------------------------------------------------------------------------------
def difflog_events(im, prev_im, neg_thresh=0.2, pos_thresh=0.2, eps=1e-5, inputs_are_log=False):
    im, prev_im = np.asarray(im, dtype=np.float64), np.asarray(prev_im, dtype=np.float64)
    difflog = im - prev_im if inputs_are_log else np.log(im + eps) - np.log(prev_im + eps)
    events = np.zeros_like(difflog)
    if np.abs(difflog).max() < max(pos_thresh, neg_thresh):
        return events
    pos, neg = np.where(difflog > 0.0), np.where(difflog < 0.0)
    events[pos] = (difflog[pos] // pos_thresh) * pos_thresh
    events[neg] = (difflog[neg] // -neg_thresh) * -neg_thresh
    return events


def normalize_event_frames(ev, rescale_evs=-1.0, evs_min_cutoff=None):
    """learner/dataloading.py:508-533 for one trajectory [T,H,W], with torch like the reference."""
    import torch
    ev = torch.as_tensor(np.asarray(ev, dtype=np.float32)).clone()
    if rescale_evs > 0.0:
        ev = torch.clamp(ev / rescale_evs, -1.0, 1.0)
    elif rescale_evs == -1.0:
        maxvals = torch.quantile(torch.abs(ev).view(ev.shape[0], -1), 0.97, dim=1).view(ev.shape[0], 1, 1)
        ev = torch.clamp(ev / maxvals, -1.0, 1.0)
    if evs_min_cutoff is not None:
        ev[ev.abs() < evs_min_cutoff] = 0.0
    return ev.numpy()


# ---------------------------------------------------------------------------------------------
# "next" row N3: event-frame / depth rectification (utils/calibration_tools/rectify_bag.py:91-138,
# evfly_ros/run.py:339-340). The arithmetic lives in a third-party dependency, OpenCV's cv2.remap (4.13 in this
# image; imgproc/src/imgwarp.cpp: remap() fixed-point map conversion + remapBicubic<Cast<float,float>,float,1>,
# BORDER_CONSTANT 0); restated below and pinned bit-for-bit against cv2.remap and against the reference's own
# Aligner run on its shipped calibration (tests/golden/make_golden_remap.py).
# ---------------------------------------------------------------------------------------------
def cubic_table() -> np.ndarray:
    """OpenCV interpolateCubic (A = -0.75) at x = i/32, float32 [32][4]; every entry is exact in float32."""
    f32 = np.float32
    A = f32(-0.75)
    tab = np.zeros((32, 4), f32)
    for i in range(32):
        x = f32(i) * f32(1.0 / 32)
        x1, y = x + f32(1), f32(1) - x
        tab[i, 0] = ((A * x1 - f32(5) * A) * x1 + f32(8) * A) * x1 - f32(4) * A
        tab[i, 1] = ((A + f32(2)) * x - (A + f32(3))) * x * x + f32(1)
        tab[i, 2] = ((A + f32(2)) * y - (A + f32(3))) * y * y + f32(1)
        tab[i, 3] = f32(1) - tab[i, 0] - tab[i, 1] - tab[i, 2]
    return tab


def remap_bicubic(src, mapx, mapy) -> np.ndarray:
    """cv2.remap(src float32 [H,W], mapx, mapy float32 [OH,OW], INTER_CUBIC) with the default constant-0 border.
    Coordinates are rounded to 1/32 pixel (round-half-even of map*32), the 16 weights are float32 products of the
    two 1-D table rows; interior pixels sum row by row (((a+b)+c)+d per row, rows added in order), pixels whose
    4x4 window leaves the image accumulate the in-range taps one by one from 0. No fused multiply-add."""
    f32 = np.float32
    src = np.ascontiguousarray(src, dtype=f32)
    H, W = src.shape
    mapx, mapy = np.asarray(mapx, dtype=f32), np.asarray(mapy, dtype=f32)
    tab = cubic_table()
    sx = np.rint(mapx * f32(32)).astype(np.int64)
    sy = np.rint(mapy * f32(32)).astype(np.int64)
    fx, fy = sx & 31, sy & 31
    ix = np.clip(sx >> 5, -32768, 32767) - 1
    iy = np.clip(sy >> 5, -32768, 32767) - 1
    wx, wy = tab[fx], tab[fy]                                   # [...,4]
    interior = (ix >= 0) & (ix < max(W - 3, 0)) & (iy >= 0) & (iy < max(H - 3, 0))
    fast = None
    slow = np.zeros(mapx.shape, f32)
    for k1 in range(4):
        yy = iy + k1
        yok = (yy >= 0) & (yy < H)
        row = None
        for k2 in range(4):
            xx = ix + k2
            ok = yok & (xx >= 0) & (xx < W)
            v = np.where(ok, src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], f32(0))
            term = (v * (wy[..., k1] * wx[..., k2]).astype(f32)).astype(f32)
            row = term if row is None else (row + term).astype(f32)
            slow = np.where(ok, (slow + term).astype(f32), slow)
        fast = row if fast is None else (fast + row).astype(f32)
    return np.where(interior, fast, slow).astype(f32)


def remap_img(img, maps, flip=False, rotate=False) -> np.ndarray:
    """rectify_bag.py:91-98."""
    img = np.asarray(img, dtype=np.float32)
    if flip:
        img = img[:, ::-1]
    out = remap_bicubic(img, maps[0], maps[1])
    return out[::-1, ::-1].copy() if rotate else out          # cv2.rotate(ROTATE_180)


def remap_events(events: dict, maps, rotate: bool, shape):
    """rectify_bag.py:101-116: per-event lookup of the inverse maps + in-frame mask (coordinates stay float)."""
    mx, my = maps
    x, y = mx[events["y"], events["x"]], my[events["y"], events["x"]]
    tw, th = shape
    if rotate:
        x, y = tw - 1 - x, th - 1 - y
    m = (x >= 0) & (x <= tw - 1) & (y >= 0) & (y <= th - 1)
    return {"x": x[m], "y": y[m], "t": events["t"][m], "p": events["p"][m]}
