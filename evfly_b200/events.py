"""Event record format and the host<->device plumbing around libevfly_b200's L1 kernels.

The record is the in-memory layout of dv_ros_msgs::Event / prophesee_event_msgs::Event
(dv_ros_msgs/msg/Event.msg:2-5), 16 bytes; on the device it is a uint8 tensor [n,16].
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

EVENT_DTYPE = np.dtype([
    ("x", "<u2"), ("y", "<u2"), ("ts_sec", "<u4"), ("ts_nsec", "<u4"),
    ("polarity", "u1"), ("pad", "u1", (3,)),
])
assert EVENT_DTYPE.itemsize == 16

NS = 1_000_000_000


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.EvflyError("evfly_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def make_records(x, y, t_ns, p) -> np.ndarray:
    """Host-side constructor of EVENT_DTYPE records from integer arrays (p: 1 = +, 0 = -)."""
    x = np.asarray(x)
    n = x.shape[0]
    rec = np.zeros(n, dtype=EVENT_DTYPE)
    rec["x"] = x
    rec["y"] = np.asarray(y)
    t = np.asarray(t_ns, dtype=np.int64)
    rec["ts_sec"] = t // NS
    rec["ts_nsec"] = t % NS
    rec["polarity"] = np.asarray(p)
    return rec


def records_time_ns(rec: np.ndarray) -> np.ndarray:
    return rec["ts_sec"].astype(np.int64) * NS + rec["ts_nsec"].astype(np.int64)


def to_device(records, device=None, pinned: torch.Tensor | None = None) -> torch.Tensor:
    """EVENT_DTYPE array / bytes / uint8 tensor -> uint8 CUDA tensor [n,16].

    `pinned`: optional reusable pinned staging buffer (uint8 [cap,16]); the host copy into it and
    the asynchronous H2D copy out of it are what a ROS callback would do with msg->events.
    """
    dev = _device(device)
    if isinstance(records, torch.Tensor):
        t = records
        if t.dtype != torch.uint8 or t.dim() != 2 or t.shape[1] != 16:
            raise ValueError("event record tensors must be uint8 [n,16]")
        return t.to(dev, non_blocking=True).contiguous()
    if isinstance(records, (bytes, bytearray, memoryview)):
        records = np.frombuffer(records, dtype=EVENT_DTYPE)
    records = np.ascontiguousarray(records)
    if records.dtype != EVENT_DTYPE:
        raise ValueError(f"expected EVENT_DTYPE records, got {records.dtype}")
    host = torch.from_numpy(records.view(np.uint8).reshape(-1, 16))
    if pinned is not None and pinned.shape[0] >= host.shape[0]:
        stage = pinned[: host.shape[0]]
        stage.copy_(host)
        return stage.to(dev, non_blocking=True)
    return host.to(dev)


class L1:
    """Thin, typed wrappers of the L1 entry points; all tensors live on the device."""

    @staticmethod
    def pack_rows_f64(rows: torch.Tensor, H: int, W: int, pol_mode: int, *, t_lo=None, t_hi=None,
                      max_events: int | None = None):
        """rows: float64 CUDA [n,4] = (t,x,y,p). Returns (records uint8 [n,16], last_kept_t)."""
        lib = _lib.load()
        n = rows.shape[0]
        out = torch.empty((n, 16), dtype=torch.uint8, device=rows.device)
        last = torch.full((1,), float("nan"), dtype=torch.float64, device=rows.device)
        use_time = int(t_lo is not None and t_hi is not None and max_events is None)
        ws = None
        if max_events is not None:
            ws = torch.empty((n // 1024 + 3,), dtype=torch.int64, device=rows.device)
        _lib.check(lib.evfly_pack_events_f64(
            _lib.ptr(rows), n, H, W, pol_mode, use_time,
            float(t_lo) if t_lo is not None else 0.0, float(t_hi) if t_hi is not None else 0.0,
            -1 if max_events is None else int(max_events),
            _lib.ptr(out), _lib.ptr(last), _lib.ptr(ws), _lib.stream_ptr()), "evfly_pack_events_f64")
        return out, last

    @staticmethod
    def pack_soa(x, y, t_ns, p, H: int, W: int, pol_mode: int) -> torch.Tensor:
        lib = _lib.load()
        x, y, t_ns, p = (v.to(torch.int64).contiguous() for v in (x, y, t_ns, p))
        n = x.shape[0]
        out = torch.empty((n, 16), dtype=torch.uint8, device=x.device)
        _lib.check(lib.evfly_pack_events_soa(_lib.ptr(x), _lib.ptr(y), _lib.ptr(t_ns), _lib.ptr(p), n,
                                             H, W, pol_mode, _lib.ptr(out), _lib.stream_ptr()),
                   "evfly_pack_events_soa")
        return out

    _binned_ws: dict = {}

    @staticmethod
    def accumulate_counts(records: torch.Tensor, H: int, W: int, out: torch.Tensor | None = None, algo: str = "red"):
        """counts[pol][y][x] += #events. algo: 'red' (one L2 reduction per event, the default: 93 us for 10 M events)
        or 'binned' (spatial binning + shared-memory histograms; measured 109 us, kept as a documented experiment --
        DESIGN.md section 8)."""
        lib = _lib.load()
        if out is None:
            out = torch.zeros((2, H, W), dtype=torch.int32, device=records.device)
        n = records.shape[0]
        if algo == "binned":
            need = lib.evfly_accumulate_counts_binned_workspace_bytes(n, H, W)
            key = (records.device, torch.cuda.current_stream().cuda_stream)
            ws = L1._binned_ws.get(key)
            if ws is None or ws.numel() < need:
                ws = torch.zeros((need,), dtype=torch.uint8, device=records.device)      # cursors zero; left zero by the kernel
                L1._binned_ws[key] = ws
            _lib.check(lib.evfly_accumulate_counts_binned(_lib.ptr(records), n, H, W, _lib.ptr(out), _lib.ptr(ws),
                                                          _lib.stream_ptr()), "evfly_accumulate_counts_binned")
            return out
        _lib.check(lib.evfly_accumulate_counts(_lib.ptr(records), n, H, W,
                                               _lib.ptr(out), _lib.stream_ptr()), "evfly_accumulate_counts")
        return out

    @staticmethod
    def counts_to_frame_f64(counts: torch.Tensor, pos_thresh=0.2, neg_thresh=0.2) -> torch.Tensor:
        lib = _lib.load()
        _, H, W = counts.shape
        out = torch.empty((H, W), dtype=torch.float64, device=counts.device)
        _lib.check(lib.evfly_counts_to_frame_f64(_lib.ptr(counts), H, W, float(pos_thresh),
                                                 float(neg_thresh), _lib.ptr(out), _lib.stream_ptr()),
                   "evfly_counts_to_frame_f64")
        return out

    @staticmethod
    def counts_to_u8(counts: torch.Tensor, mode: int, state: torch.Tensor | None = None,
                     flagged_cap: int = 4096):
        """Returns (frame u8 [H,W], flagged int32 [cap], n_flagged int32 [1])."""
        lib = _lib.load()
        _, H, W = counts.shape
        out = torch.empty((H, W), dtype=torch.uint8, device=counts.device)
        flagged = torch.empty((flagged_cap,), dtype=torch.int32, device=counts.device)
        nfl = torch.zeros((1,), dtype=torch.int32, device=counts.device)
        _lib.check(lib.evfly_counts_to_u8(_lib.ptr(counts), H, W, mode, _lib.ptr(state), _lib.ptr(out),
                                          _lib.ptr(flagged), flagged_cap, _lib.ptr(nfl),
                                          _lib.stream_ptr()), "evfly_counts_to_u8")
        return out, flagged, nfl

    @staticmethod
    def u8_saturate_replay(records, H, W, state, flagged, n_flagged: int, frame):
        lib = _lib.load()
        _lib.check(lib.evfly_u8_saturate_replay(_lib.ptr(records), records.shape[0], H, W,
                                                _lib.ptr(state), _lib.ptr(flagged), int(n_flagged),
                                                _lib.ptr(frame), _lib.stream_ptr()),
                   "evfly_u8_saturate_replay")
        return frame

    @staticmethod
    def voxel_workspace(H: int, W: int, B: int, device) -> torch.Tensor:
        nbytes = _lib.load().evfly_voxel_workspace_bytes(H, W, B)
        return torch.zeros((nbytes // 4,), dtype=torch.float32, device=device)

    @staticmethod
    def voxelize_window(records, H, W, B, t0_ns, t1_ns, *, counts=None, voxel=None, ws=None,
                        algo=1, want_counts=True, want_voxel=True):
        """counts / voxel must be zero on entry when given; returns (counts|None, voxel|None)."""
        lib = _lib.load()
        dev = records.device
        if want_counts and counts is None:
            counts = torch.zeros((2, H, W), dtype=torch.int32, device=dev)
        if want_voxel and voxel is None:
            voxel = torch.zeros((B, H, W), dtype=torch.float32, device=dev)
        if algo == 1 and ws is None:
            ws = L1.voxel_workspace(H, W, B, dev)
        _lib.check(lib.evfly_voxelize_window(_lib.ptr(records), records.shape[0], H, W, B, int(t0_ns),
                                             int(t1_ns), _lib.ptr(counts), _lib.ptr(voxel),
                                             _lib.ptr(ws), algo, _lib.stream_ptr()),
                   "evfly_voxelize_window")
        return counts, voxel

    @staticmethod
    def accumulate_windows(records, edges_ns: torch.Tensor, H, W, B=None, *, sorted_by_time=True,
                           counts=None, voxel=None):
        """edges_ns int64 CUDA [T+1]. Returns (counts int32 [T,2,H,W], voxel fp32 [T,B,H,W]|None)."""
        lib = _lib.load()
        dev = records.device
        T = edges_ns.shape[0] - 1
        if counts is None:
            counts = torch.empty((T, 2, H, W), dtype=torch.int32, device=dev)
        if B is not None and voxel is None:
            voxel = torch.empty((T, B, H, W), dtype=torch.float32, device=dev)
        rng = torch.empty((T + 1,), dtype=torch.int64, device=dev)
        _lib.check(lib.evfly_accumulate_windows(_lib.ptr(records), records.shape[0], _lib.ptr(edges_ns),
                                                T, H, W, 0 if B is None else B, _lib.ptr(counts),
                                                _lib.ptr(voxel), int(bool(sorted_by_time)),
                                                _lib.ptr(rng), _lib.stream_ptr()),
                   "evfly_accumulate_windows")
        return counts, voxel
