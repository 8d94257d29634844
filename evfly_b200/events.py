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

# 8-byte wire record (include/evfly_b200.h evfly_event8): offset from the window's first edge instead of the
# absolute time; x = 0xffff marks a skipped record
EVENT8_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("dt_pol", "<u4")])
assert EVENT8_DTYPE.itemsize == 8

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


def pack_ev8_host(records: np.ndarray, edges_ns: np.ndarray):
    """Host-side packer of the wire format (numpy): EVENT_DTYPE records of ONE time-sorted stream + window edges
    int64 [T+1] -> (EVENT8_DTYPE records [n], win_offsets int64 [T+1]). Same result as L1.pack_ev8 on the device."""
    t = records_time_ns(records)
    edges_ns = np.asarray(edges_ns, dtype=np.int64)
    offsets = np.searchsorted(t, edges_ns, side="left").astype(np.int64)
    w = np.clip(np.searchsorted(offsets, np.arange(records.shape[0]), side="right") - 1, 0, len(edges_ns) - 2)
    dt = t - edges_ns[w]
    out = np.zeros(records.shape[0], dtype=EVENT8_DTYPE)
    idx = np.arange(records.shape[0])
    ok = (idx >= offsets[0]) & (idx < offsets[-1]) & (records["polarity"] < 2) & (dt >= 0) & (dt < (1 << 31)) & (t < edges_ns[w + 1])
    out["x"] = np.where(ok, records["x"], 0xFFFF)
    out["y"] = np.where(ok, records["y"], 0xFFFF)
    out["dt_pol"] = np.where(ok, (dt.astype(np.uint64) << np.uint64(1)) | records["polarity"].astype(np.uint64), 0).astype(np.uint32)
    return out, offsets


# 4-byte wire record (include/evfly_b200.h evfly_event4): x [0,10) | y [10,19) | polarity [19] | delta [20,32) = microseconds since the
# previous record of the same window (the window's first edge for its first record). (1023, 511) is a skip record that only
# advances the time (gaps above 4095 us). Exact for streams on a 1 us grid (every DVS sensor's timestamps), H <= 511, W <= 1023.
EV4_MAX_DELTA = 4095
EV4_SKIP = 1023 | (511 << 10)
CHUNK_EVENTS = 8192        # evfly_accumulate_chunk_events(): the device sorts windows in chunks of this many records


def pack_ev4_host(records: np.ndarray, edges_ns: np.ndarray):
    """Host-side packer of the 4-byte wire format: EVENT_DTYPE records of ONE time-sorted stream + window edges int64 [T+1] ->
    (uint32 records [m], win_offsets int64 [T+1], chunk_base_us uint32 [sum_w ceil(n_w / CHUNK_EVENTS)]) or None when the stream
    cannot be represented exactly (times or edges off the 1 us grid, not sorted, coordinates >= (1023, 511)). Events that can
    never count (outside the windows, polarity >= 2) are dropped here; chunk_base_us[c] is the offset from the window's first edge of the last record
    before chunk c of that window."""
    t = records_time_ns(records)
    edges_ns = np.asarray(edges_ns, dtype=np.int64)
    T = len(edges_ns) - 1
    if T < 1 or (edges_ns % 1000).any() or (np.diff(edges_ns) < 0).any() or (t.size > 1 and (np.diff(t) < 0).any()):
        return None
    offsets = np.searchsorted(t, edges_ns, side="left")
    lo, hi = int(offsets[0]), int(offsets[-1])
    r, tt = records[lo:hi], t[lo:hi]
    w = np.clip(np.searchsorted(offsets, np.arange(lo, hi), side="right") - 1, 0, T - 1)
    keep = r["polarity"] < 2
    r, tt, w = r[keep], tt[keep], w[keep]
    if r.size and (int(r["x"].max()) >= 1023 or int(r["y"].max()) >= 511):
        return None                      # coordinates beyond the record's 10 + 9 bits (a larger sensor): the 8-byte record takes them
    dt = tt - edges_ns[w]
    if (dt % 1000).any():
        return None
    dt_us = dt // 1000
    first = np.ones(len(r), dtype=bool)
    first[1:] = w[1:] != w[:-1]
    prev = np.zeros(len(r), dtype=np.int64)
    prev[1:] = dt_us[:-1]
    prev[first] = 0
    delta = dt_us - prev
    n_esc = np.maximum(delta - 1, 0) // EV4_MAX_DELTA
    rem = delta - n_esc * EV4_MAX_DELTA
    per = 1 + n_esc
    pos = np.cumsum(per) - 1
    total = int(per.sum())
    out = np.full(total, EV4_SKIP | (EV4_MAX_DELTA << 20), dtype=np.uint32)
    out[pos] = (r["x"].astype(np.uint32) | (r["y"].astype(np.uint32) << 10) | (r["polarity"].astype(np.uint32) << 19) | (rem.astype(np.uint32) << 20))
    cnt = np.bincount(w, weights=per, minlength=T).astype(np.int64)
    offs = np.zeros(T + 1, dtype=np.int64)
    np.cumsum(cnt, out=offs[1:])
    cum = np.cumsum((out >> 20).astype(np.int64))                     # running time over the whole packed stream
    bases = []
    for k in range(T):
        nchunk = -(-int(cnt[k]) // CHUNK_EVENTS)
        if nchunk == 0:
            continue
        start = cum[offs[k] - 1] if offs[k] > 0 else 0
        idx = offs[k] + np.arange(1, nchunk, dtype=np.int64) * CHUNK_EVENTS - 1
        bases.append(np.concatenate([[0], cum[idx] - start]))
    chunk_base = np.concatenate(bases).astype(np.uint32) if bases else np.zeros(0, dtype=np.uint32)
    return out, offs, chunk_base


class WireBatch:
    """A batch of n_traj trajectories of T windows each in a wire format, ready for L1.accumulate_windows_wire /
    PerceptionPipeline: records of all trajectories laid end to end (uint8 [n, record_bytes]; a pinned HOST tensor for
    the feeder, a device tensor for the pipeline) plus the small per-window tables (device int64 / int32): event index
    range, time range, and the output slot t*n_traj + s (time-major frames: the model advances the trajectories
    together). record_bytes = 8: evfly_event8; 4: evfly_event4 (delta-coded time, chunk_base = the per-chunk time table)."""

    def __init__(self, records, win_offsets, win_t0, win_t1, out_slot, n_traj, T, chunk_base=None):
        self.records, self.win_offsets, self.win_t0, self.win_t1, self.out_slot = records, win_offsets, win_t0, win_t1, out_slot
        self.n_traj, self.T = int(n_traj), int(T)
        self.chunk_base = chunk_base
        self.record_bytes = int(records.shape[1])

    @property
    def n_windows(self):
        return self.n_traj * self.T

    @staticmethod
    def from_streams(streams, device, pin=True, fmt="auto"):
        """streams: list of (EVENT_DTYPE records, edges int64 [T+1]) of equal T (host numpy). Packs on the host.
        fmt: 8 = evfly_event8, 4 = evfly_event4 (raises if a stream is not on the 1 us grid), "auto" = 4 when every stream
        can be represented exactly in it, else 8."""
        T = len(streams[0][1]) - 1
        assert all(len(e) - 1 == T for _, e in streams), "trajectories must have the same number of windows"
        n = len(streams)
        packed4 = None
        if fmt in ("auto", 4):
            packed4 = [pack_ev4_host(r, e) for r, e in streams]
            if any(p is None for p in packed4):
                if fmt == 4:
                    raise ValueError("a stream cannot be represented in the 4-byte wire format (times off the 1 us grid, or not sorted)")
                packed4 = None
        recs, offs, t0, t1, cb, base = [], [], [], [], [], 0
        for i, (r, e) in enumerate(streams):
            if packed4 is not None:
                r_p, o, c = packed4[i]
                cb.append(c)
            else:
                r_p, o = pack_ev8_host(r, e)
            recs.append(r_p)
            # window (s,t) = [base + o[t], next start): records of a trajectory that lie outside its windows were made
            # skip records by the 8-byte packer (dropped by the 4-byte one), so a range may harmlessly include them
            offs.append(o[:-1] + base)
            t0.append(np.asarray(e[:-1], np.int64)); t1.append(np.asarray(e[1:], np.int64))
            base += r_p.shape[0]
        offs.append(np.array([base], dtype=np.int64))
        rb = 4 if packed4 is not None else 8
        host = torch.from_numpy(np.concatenate(recs).view(np.uint8).reshape(-1, rb))
        if pin:
            host = host.pin_memory()
        slot = (np.arange(T, dtype=np.int32)[None, :] * n + np.arange(n, dtype=np.int32)[:, None]).reshape(-1)
        dev = torch.device(device)
        chunk_base = torch.from_numpy(np.concatenate(cb).view(np.int32)).to(dev) if packed4 is not None else None
        return WireBatch(host, torch.from_numpy(np.concatenate(offs)).to(dev), torch.from_numpy(np.concatenate(t0)).to(dev),
                         torch.from_numpy(np.concatenate(t1)).to(dev), torch.from_numpy(slot).to(dev), n, T, chunk_base)

    def on_device(self, device_records):
        return WireBatch(device_records, self.win_offsets, self.win_t0, self.win_t1, self.out_slot, self.n_traj, self.T, self.chunk_base)


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

    _sorted_ws: dict = {}

    @staticmethod
    def sorted_workspace(n, n_windows, H, W, B, device) -> torch.Tensor:
        """Scratch of the shared-memory-tile path, cached per (device, stream) and grown on demand."""
        need = _lib.load().evfly_accumulate_sorted_workspace_bytes(int(n), int(n_windows), H, W, 0 if B is None else B)
        if need <= 0:
            raise _lib.EvflyError(f"accumulate (tiled): unsupported shape H={H} W={W} B={B}")
        key = (torch.device(device), torch.cuda.current_stream().cuda_stream)
        ws = L1._sorted_ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty((int(need * 1.25) + 4096,), dtype=torch.uint8, device=device)
            L1._sorted_ws[key] = ws
        return ws

    @staticmethod
    def pack_ev8(records, edges_ns: torch.Tensor):
        """Device packer: canonical records [n,16] of one time-sorted stream -> (wire records uint8 [n,8], win_offsets int64 [T+1])."""
        lib = _lib.load()
        n, T = records.shape[0], edges_ns.shape[0] - 1
        out = torch.empty((n, 8), dtype=torch.uint8, device=records.device)
        offs = torch.empty((T + 1,), dtype=torch.int64, device=records.device)
        _lib.check(lib.evfly_pack_events_ev8(_lib.ptr(records), n, _lib.ptr(edges_ns), T, _lib.ptr(out), _lib.ptr(offs), _lib.stream_ptr()),
                   "evfly_pack_events_ev8")
        return out, offs

    @staticmethod
    def accumulate_windows_ev8(records8, win_offsets, win_t0, win_t1, H, W, B=None, *, out_slot=None, n_slots=None, counts=None, voxel=None):
        """Wire records uint8 [n,8] (any number of streams end to end) + per-window tables on the device ->
        (counts int32 [n_slots,2,H,W], voxel fp32 [n_slots,B,H,W] | None); window w lands in slot out_slot[w] (default w)."""
        lib = _lib.load()
        dev = records8.device
        nw = win_t0.shape[0]
        n_slots = nw if n_slots is None else n_slots
        if counts is None:
            counts = torch.empty((n_slots, 2, H, W), dtype=torch.int32, device=dev)
        if B is not None and voxel is None:
            voxel = torch.empty((n_slots, B, H, W), dtype=torch.float32, device=dev)
        ws = L1.sorted_workspace(records8.shape[0], nw, H, W, B, dev)
        _lib.check(lib.evfly_accumulate_windows_ev8(_lib.ptr(records8), records8.shape[0], _lib.ptr(win_offsets), _lib.ptr(win_t0), _lib.ptr(win_t1),
                                                    _lib.ptr(out_slot), nw, H, W, 0 if B is None else B, _lib.ptr(counts), _lib.ptr(voxel),
                                                    _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "evfly_accumulate_windows_ev8")
        return counts, voxel

    @staticmethod
    def accumulate_windows_ev4(records4, win_offsets, win_t0, win_t1, chunk_base, H, W, B=None, *, out_slot=None, n_slots=None, counts=None, voxel=None):
        """4-byte wire records uint8 [n,4] + per-window tables + the per-chunk time table (pack_ev4_host) on the device -> as
        accumulate_windows_ev8. Same frames as the 8-byte format of the same stream (counts bit-exact; identical event times)."""
        lib = _lib.load()
        dev = records4.device
        nw = win_t0.shape[0]
        n_slots = nw if n_slots is None else n_slots
        if counts is None:
            counts = torch.empty((n_slots, 2, H, W), dtype=torch.int32, device=dev)
        if B is not None and voxel is None:
            voxel = torch.empty((n_slots, B, H, W), dtype=torch.float32, device=dev)
        ws = L1.sorted_workspace(records4.shape[0], nw, H, W, B, dev)
        _lib.check(lib.evfly_accumulate_windows_ev4(_lib.ptr(records4), records4.shape[0], _lib.ptr(win_offsets), _lib.ptr(win_t0), _lib.ptr(win_t1),
                                                    _lib.ptr(out_slot), _lib.ptr(chunk_base), nw, H, W, 0 if B is None else B, _lib.ptr(counts),
                                                    _lib.ptr(voxel), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "evfly_accumulate_windows_ev4")
        return counts, voxel

    @staticmethod
    def accumulate_windows_wire(wb, H, W, B=None, *, counts=None, voxel=None):
        """A WireBatch (records on the device) in either wire format -> frames in its time-major slots."""
        kw = dict(out_slot=wb.out_slot, n_slots=wb.n_traj * wb.T, counts=counts, voxel=voxel)
        if wb.record_bytes == 4:
            return L1.accumulate_windows_ev4(wb.records, wb.win_offsets, wb.win_t0, wb.win_t1, wb.chunk_base, H, W, B, **kw)
        return L1.accumulate_windows_ev8(wb.records, wb.win_offsets, wb.win_t0, wb.win_t1, H, W, B, **kw)

    @staticmethod
    def accumulate_windows(records, edges_ns: torch.Tensor, H, W, B=None, *, sorted_by_time=True,
                           counts=None, voxel=None, algo="auto", slot_stride=1, slot_offset=0):
        """edges_ns int64 CUDA [T+1]. Returns (counts int32 [T,2,H,W], voxel fp32 [T,B,H,W]|None).
        algo: 'tiles' = shared-memory histogram tiles (time-sorted streams; window w goes to frame slot
        w*slot_stride + slot_offset of the given buffers), 'scatter' = L2 reductions (any order), 'auto' = tiles
        for sorted streams."""
        lib = _lib.load()
        dev = records.device
        T = edges_ns.shape[0] - 1
        if counts is None:
            counts = torch.empty((T, 2, H, W), dtype=torch.int32, device=dev)
        if B is not None and voxel is None:
            voxel = torch.empty((T, B, H, W), dtype=torch.float32, device=dev)
        if algo == "auto":
            algo = "tiles" if sorted_by_time else "scatter"
        if algo == "tiles":
            if not sorted_by_time:
                raise _lib.EvflyError("accumulate_windows(algo='tiles') needs a time-sorted stream")
            ws = L1.sorted_workspace(records.shape[0], T, H, W, B, dev)
            _lib.check(lib.evfly_accumulate_windows_sorted(_lib.ptr(records), records.shape[0], _lib.ptr(edges_ns), T, H, W, 0 if B is None else B,
                                                           _lib.ptr(counts), _lib.ptr(voxel), int(slot_stride), int(slot_offset), _lib.ptr(ws),
                                                           ws.numel(), _lib.stream_ptr()), "evfly_accumulate_windows_sorted")
            return counts, voxel
        if slot_stride != 1 or slot_offset != 0:
            raise _lib.EvflyError("accumulate_windows(algo='scatter') writes consecutive frames only")
        rng = torch.empty((T + 1,), dtype=torch.int64, device=dev)
        _lib.check(lib.evfly_accumulate_windows(_lib.ptr(records), records.shape[0], _lib.ptr(edges_ns),
                                                T, H, W, 0 if B is None else B, _lib.ptr(counts),
                                                _lib.ptr(voxel), int(bool(sorted_by_time)),
                                                _lib.ptr(rng), _lib.stream_ptr()),
                   "evfly_accumulate_windows")
        return counts, voxel
