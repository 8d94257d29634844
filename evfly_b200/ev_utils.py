"""Drop-in for the event-binning functions of evfly's `utils/ev_utils.py` and the per-window
slicing of `utils/to_events.py`, computed by libevfly_b200's scatter kernels.

`form_eventframe` keeps the reference's signature, argument meaning, return types (numpy
float64, plus times1 in the timed modes) and error behaviour (utils/ev_utils.py:113-161).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .events import L1, _device


def _rows_to_device(view_events) -> torch.Tensor:
    """[n,>=4] array with columns (t, x, y, ..., p) -> float64 CUDA [n,4] = (t,x,y,p)."""
    dev = _device()
    if isinstance(view_events, torch.Tensor):
        ev = view_events
    else:
        ev = torch.from_numpy(np.ascontiguousarray(np.asarray(view_events)))
    if ev.dim() != 2 or ev.shape[1] < 4:
        raise ValueError("view_events must be [n,4] with columns (t, x, y, p)")
    ev = ev.to(dev, non_blocking=True)
    if ev.shape[1] != 4:  # the reference reads columns 0,1,2 and -1
        ev = torch.cat([ev[:, :3], ev[:, -1:]], dim=1)
    return ev.to(torch.float64).contiguous()


def form_eventframe(view_events, H, W, times0=None, times1=None, N=None, device='cpu',
                    is_half_res=False, pos_thresh=0.2, neg_thresh=0.2, all_events=False):
    """Signed event-count frame, `pos_thresh*#pos - neg_thresh*#neg` per pixel (float64 [H,W]).

    Mirrors utils/ev_utils.py:113-161:
      all_events=True : every row; positive p>0, negative p==0; returns frame.
      all_events=False: rows with times0*1e9 <= t < times1[0]*1e9 (or, if times1 is None, the
                        first N rows with t >= times0*1e9); positive p>0, negative p<0;
                        returns (frame, times1).
    `device` and `is_half_res` are accepted and unused, as in the reference.
    """
    if not all_events:
        if len(view_events) == 0:
            return np.zeros((H, W)), times0
        if times0 is None:
            print('times0 argument is None but it must be given to establilsh a starting point for the events slicing!')
            raise SystemExit
        rows = _rows_to_device(view_events)
        t_lo = float(times0 * 1e9)
        if times1 is not None:
            t_hi = float(times1[0] * 1e9)
            rec, _ = L1.pack_rows_f64(rows, H, W, _lib.NEG_IS_NEGATIVE, t_lo=t_lo, t_hi=t_hi)
        elif N is not None:
            print(f'times0: {times0}')
            rec, last_t = L1.pack_rows_f64(rows, H, W, _lib.NEG_IS_NEGATIVE, t_lo=t_lo, max_events=int(N))
            last_t = float(last_t.item())
            if last_t != last_t:  # no row passed the mask: the reference indexes [-1] of an empty array
                raise IndexError("index -1 is out of bounds for axis 0 with size 0")
            times1 = (last_t + 1) / 1e9
        else:
            raise ValueError("form_eventframe() requires either times1 or N to be not None")
        counts = L1.accumulate_counts(rec, H, W)
        frame = L1.counts_to_frame_f64(counts, pos_thresh, neg_thresh)
        return frame.cpu().numpy(), times1

    if len(view_events) == 0:
        return np.zeros((H, W))
    rows = _rows_to_device(view_events)
    rec, _ = L1.pack_rows_f64(rows, H, W, _lib.NEG_IS_ZERO)
    counts = L1.accumulate_counts(rec, H, W)
    return L1.counts_to_frame_f64(counts, pos_thresh, neg_thresh).cpu().numpy()


def form_eventframes_sliced(events: dict, t_edges_ns, H, W, pos_thresh=0.2, neg_thresh=0.2,
                            return_counts=False):
    """All T windows of one continuous stream in one pass (utils/to_events.py:400-411 rescans the
    stream once per window). `events` is the Vid2E dict of 1-D tensors 'x','y','t' (ns),'p' (+-1);
    window i is [t_edges_ns[i], t_edges_ns[i+1]) -- the reference's t_start/t_end.
    Returns float64 numpy [T,H,W] (frames[i] = frame.T of the reference)."""
    dev = _device()
    get = lambda k: torch.as_tensor(events[k]).to(dev)
    rec = L1.pack_soa(get('x'), get('y'), get('t'), get('p'), H, W, _lib.NEG_IS_NEGATIVE)
    edges_f = np.asarray(t_edges_ns, dtype=np.float64)
    # t is an integer number of ns: t >= e  <=>  t >= ceil(e)
    edges = torch.from_numpy(np.ceil(edges_f).astype(np.int64)).to(dev)
    t = get('t')
    sorted_by_time = bool((t[1:] >= t[:-1]).all().item()) if t.numel() > 1 else True
    counts, _ = L1.accumulate_windows(rec, edges, H, W, None, sorted_by_time=sorted_by_time)
    if return_counts:
        return counts
    T = counts.shape[0]
    frames = torch.empty((T, H, W), dtype=torch.float64, device=dev)
    for i in range(T):
        frames[i] = L1.counts_to_frame_f64(counts[i], pos_thresh, neg_thresh)
    return frames.cpu().numpy()


def form_voxelgrid(records, H, W, t0_ns, t1_ns, num_bins=5, algo=1):
    """Temporal-bilinear voxel grid of one window of packed records (uint8 CUDA [n,16]).
    Build-defined (the reference has none):
        tau = (B-1)(t-t0)/(t1-t0);  V[b,y,x] = sum_i pol_i max(0, 1-|b-tau_i|),  pol = +-1.
    Returns (counts int32 [2,H,W], voxel fp32 [B,H,W]) on the device."""
    return L1.voxelize_window(records, H, W, num_bins, t0_ns, t1_ns, algo=algo)
