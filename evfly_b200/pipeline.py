"""The hot path end to end on one GPU: packed event records of consecutive windows ->
count frames (+ temporal voxel grid) -> decode/crop -> 97th-percentile scale + clip ->
OrigUNet_w_VITFLY_ViTLSTM forward with carried recurrent state.

Mirrors what evfly_ros/run.py does per timer tick (evs_process :330-364 + run_model :245-282)
and what learner/evaluation_tools.py does per trajectory (:62-66), with the C++ node's
accumulation (evfly_ros/src/node.cpp) folded in front.
"""
from __future__ import annotations

import torch

from . import _lib
from .events import L1


class PerceptionPipeline:
    def __init__(self, model, sensor_hw=(480, 640), model_hw=(260, 346), num_bins=5, desvel=4.0, device=None, aligner=None):
        """aligner: optional evfly_b200.calibration_tools.Aligner -- the event frame is rectified (cv2.remap bicubic)
        between decode and centre crop like evfly_ros/run.py:339-340 (--align_evframe)."""
        self.aligner = aligner
        self.model = model
        self.H, self.W = sensor_hw
        self.h, self.w = model_hw
        self.B = num_bins
        self.desvel = float(desvel)
        self.dev = next(model.parameters()).device if device is None else torch.device(device)
        if self.dev.type != "cuda":
            raise _lib.EvflyError("PerceptionPipeline needs the model on a CUDA device")
        self.state_unet = None
        self.state_vit = None

    def reset(self):
        self.state_unet = self.state_vit = None

    # ---- L1 + L2 ---------------------------------------------------------------------------------
    def frames_from_windows(self, records, edges_ns, want_voxel=True, sorted_by_time=True):
        """records uint8 [n,16] on the device, edges_ns int64 [T+1] on the device.
        Returns (frames fp32 [T,1,h,w] normalised like run.py:250-253, counts, voxel)."""
        lib = _lib.load()
        T = edges_ns.shape[0] - 1
        counts, voxel = L1.accumulate_windows(records, edges_ns, self.H, self.W, self.B if want_voxel else None,
                                              sorted_by_time=sorted_by_time)
        frames = torch.empty((T, 1, self.h, self.w), dtype=torch.float32, device=self.dev)
        self._normalise(counts, frames)
        return frames, counts, voxel

    def _normalise(self, counts, frames):
        """counts int32 [N,2,H,W] -> frames fp32 [N,1,h,w]: 0.2*(n+ - n-), optional rectification, centre crop,
        per-frame 97th-percentile scaling and clip (run.py:334-351, 250-253)."""
        lib = _lib.load()
        N = counts.shape[0]
        st = _lib.stream_ptr()
        if self.aligner is None:
            # integer counts -> normalised frame in one kernel (exact order statistics from an integer histogram)
            _lib.check(lib.evfly_counts_normalise(_lib.ptr(counts), N, self.H, self.W, self.h, self.w, 0.2, 0.97, -1.0, 1.0, 0.0,
                                                  _lib.ptr(frames), None, st), "evfly_counts_normalise")
            return
        else:
            # decode at full resolution, rectify, then centre-crop; only the cropped window of the remap is computed
            # (the maps are indexed by OUTPUT pixel)
            from .calibration_tools import remap_bicubic
            full = torch.empty((N, self.H, self.W), dtype=torch.float32, device=self.dev)
            _lib.check(lib.evfly_decode_crop(None, _lib.ptr(counts), N, self.H, self.W, self.H, self.W, 0.2,
                                             _lib.ptr(full), st), "evfly_decode_crop")
            mx, my = self.aligner.davis_window(self.h, self.w)
            remap_bicubic(full, mx, my, out=frames.view(N, self.h, self.w))
        _lib.check(lib.evfly_quantile_scale_clip(_lib.ptr(frames), N, self.h * self.w, 0.97, -1.0, 1.0, 0.0,
                                                 _lib.ptr(frames), None, st), "evfly_quantile_scale_clip")

    # ---- L3 ----------------------------------------------------------------------------------------
    def forward(self, frames, carry_state=True):
        """frames [T,1,h,w] = one sequence of T consecutive windows. Returns (vel [T,3], depth)."""
        T = frames.shape[0]
        desvel = torch.full((T, 1), self.desvel, dtype=torch.float32, device=self.dev)
        vel, (depth, _, ((hu, _), hv)) = self.model([frames, desvel, [self.state_unet, None], self.state_vit])
        if carry_state:
            self.state_unet, self.state_vit = hu, hv
        return vel, depth

    def frames_from_trajectories(self, records_list, edges_list, want_voxel=True, ready=None):
        """L1 + L2 for n trajectories of equal length T: each trajectory is accumulated from its own event stream into
        its slice of one [n,T,...] buffer, then ONE decode(+rectify)+crop and ONE quantile launch cover all n*T frames
        (the per-frame work is independent). ready: optional CUDA events, one per trajectory, that the current stream
        waits on before touching that trajectory's records (its host-to-device copy). Returns (frames [T*n,1,h,w] in
        time-major order t*n + s, counts [n,T,2,H,W], voxel [n,T,B,H,W] | None)."""
        lib = _lib.load()
        n = len(records_list)
        T = edges_list[0].shape[0] - 1
        assert all(e.shape[0] - 1 == T for e in edges_list), "trajectories must have the same number of windows"
        counts = torch.empty((n, T, 2, self.H, self.W), dtype=torch.int32, device=self.dev)
        voxel = torch.empty((n, T, self.B, self.H, self.W), dtype=torch.float32, device=self.dev) if want_voxel else None
        for s, (rec, edges) in enumerate(zip(records_list, edges_list)):
            if ready is not None:
                torch.cuda.current_stream().wait_event(ready[s])
            L1.accumulate_windows(rec, edges, self.H, self.W, self.B if want_voxel else None, counts=counts[s],
                                  voxel=None if voxel is None else voxel[s])
        frames = torch.empty((n, T, 1, self.h, self.w), dtype=torch.float32, device=self.dev)
        self._normalise(counts.view(n * T, 2, self.H, self.W), frames.view(n * T, 1, self.h, self.w))
        return frames.transpose(0, 1).reshape(T * n, 1, self.h, self.w), counts, voxel

    def run_trajectories(self, records_list, edges_list, want_voxel=True, ready=None):
        """Config 4: several independent trajectories of equal length T on one GPU. Each trajectory is accumulated
        on its own (its events are its own stream); the model then advances all of them together, frames in
        time-major order, so the recurrent scans run n_traj-wide. Fresh state. Returns (vel [n_traj,T,3], depth
        [n_traj,T,1,h,w])."""
        n = len(records_list)
        tm = self.frames_from_trajectories(records_list, edges_list, want_voxel, ready=ready)[0]
        T = tm.shape[0] // n
        desvel = torch.full((T * n, 1), self.desvel, dtype=torch.float32, device=self.dev)
        vel, (depth, _, _) = self.model.forward_trajectories([tm, desvel, [None, None], None], n)
        return vel.view(T, n, 3).transpose(0, 1), depth.view(T, n, 1, self.h, self.w).transpose(0, 1)

    # ---- two-stage execution: L1+L2 of batch i+1 on a side stream while the model runs batch i -------------
    def prefetch_trajectories(self, records_list, edges_list, want_voxel=True, ready=None, done_event=None):
        """Start L1 + L2 (accumulation, normalisation) of a batch of trajectories on the pipeline's side stream and
        return a handle for run_prefetched(). The scatter is bound by L2 reductions and the model by the tensor
        pipes, so the two overlap well. Only non-persistent kernels run on the side stream (the ConvLSTM scan, which
        spins at a grid barrier, stays alone on the main stream). done_event (optional) is recorded on the side
        stream once the records have been consumed."""
        if getattr(self, "_prep_stream", None) is None:
            # default priority on purpose: with a HIGH-priority side stream its CTAs are placed ahead of the persistent
            # ConvLSTM scan's, whose resident CTAs then spin at the grid barrier for the missing ones (measured: 4.6x slower)
            self._prep_stream = torch.cuda.Stream(device=self.dev)
        main = torch.cuda.current_stream()
        ps = self._prep_stream
        ps.wait_stream(main)                      # inputs (and the allocator's blocks) are ordered after what main has queued
        with torch.cuda.stream(ps):
            tm, counts, voxel = self.frames_from_trajectories(records_list, edges_list, want_voxel, ready=ready)
            if done_event is not None:
                done_event.record(ps)
            ev = torch.cuda.Event()
            ev.record(ps)
        return (tm, counts, voxel, ev, len(records_list))

    def run_prefetched(self, handle):
        """L3 for a batch prepared by prefetch_trajectories(). Returns (vel [n,T,3], depth [n,T,1,h,w])."""
        tm, counts, voxel, ev, n = handle
        main = torch.cuda.current_stream()
        main.wait_event(ev)
        for t in (tm, counts, voxel):             # allocated on the side stream, consumed / released on main
            if t is not None:
                t.record_stream(main)
        T = tm.shape[0] // n
        desvel = torch.full((T * n, 1), self.desvel, dtype=torch.float32, device=self.dev)
        vel, (depth, _, _) = self.model.forward_trajectories([tm, desvel, [None, None], None], n)
        return vel.view(T, n, 3).transpose(0, 1), depth.view(T, n, 1, self.h, self.w).transpose(0, 1)

    def __call__(self, records, edges_ns, want_voxel=True):
        frames, counts, voxel = self.frames_from_windows(records, edges_ns, want_voxel)
        vel, depth = self.forward(frames)
        return vel, depth, counts, voxel


class TrajectoryFeeder:
    """End-to-end path for offline evaluation (learner/evaluation_tools.py:62-66 feeds trajectories one after
    another from host memory): host batches of packed event records are staged through two device buffers on a
    copy stream, so the H2D copy of batch i+1 overlaps the compute of batch i.

        feeder = TrajectoryFeeder(pipe, max_events, max_windows)
        for vel in feeder.run(batches):
            ...        # vel: pinned host tensor [n_traj, T, 3], valid until the next iteration
    A batch is (records, edges) for one trajectory or ([records...], [edges...]) for several trajectories of equal
    length that advance together (PerceptionPipeline.run_trajectories); records are pinned uint8 [n,16] host
    tensors, edges int64 device tensors [T+1]."""

    def __init__(self, pipe: "PerceptionPipeline", max_events: int, max_windows: int):
        dev = pipe.dev
        self.pipe = pipe
        self.bufs = [torch.empty((max_events, 16), dtype=torch.uint8, device=dev) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [[] for _ in range(2)]                      # per slot: one event per trajectory, its H2D copy finished
        self.free = [torch.cuda.Event() for _ in range(2)]       # compute on slot finished
        self.h_vel = [torch.empty((max_windows, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]

    @staticmethod
    def _as_lists(batch):
        recs, edges = batch
        return (list(recs), list(edges)) if isinstance(recs, (list, tuple)) else ([recs], [edges])

    def _stage(self, slot, recs):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])
            while len(self.ready[slot]) < len(recs):
                self.ready[slot].append(torch.cuda.Event())
            off = 0
            for s, r in enumerate(recs):
                self.bufs[slot][off: off + r.shape[0]].copy_(r, non_blocking=True)
                off += r.shape[0]
                self.ready[slot][s].record(self.copy_stream)      # the pipeline starts on trajectory s while s+1 is still in flight

    def run(self, batches):
        it = iter(batches)
        main = torch.cuda.current_stream()
        for e in self.free:
            e.record(main)
        cur = next(it, None)
        if cur is None:
            return
        self._stage(0, self._as_lists(cur)[0])
        slot = 0
        pending = None          # (slot, view) of the previous batch: its result is handed out one batch late, so the
                                # host is always one batch ahead of the GPU and the launch queue never drains
        handle = None           # L1+L2 of `cur`, started one iteration early on the pipeline's side stream

        def views_of(slot_, recs_):
            out_, off_ = [], 0
            for r in recs_:
                out_.append(self.bufs[slot_][off_: off_ + r.shape[0]])
                off_ += r.shape[0]
            return out_

        while cur is not None:
            nxt = next(it, None)
            recs, edges = self._as_lists(cur)
            n, T = len(recs), edges[0].shape[0] - 1
            multi = n > 1
            with torch.no_grad():
                if multi and handle is None:
                    handle = self.pipe.prefetch_trajectories(views_of(slot, recs), edges, ready=self.ready[slot][:n], done_event=self.free[slot])
                nxt_handle = None
                if nxt is not None:
                    nrecs, nedges = self._as_lists(nxt)
                    self._stage(slot ^ 1, nrecs)                          # H2D of batch i+1 ...
                    if len(nrecs) > 1:                                     # ... and its L1+L2, both overlapping the model of batch i
                        nxt_handle = self.pipe.prefetch_trajectories(views_of(slot ^ 1, nrecs), nedges, ready=self.ready[slot ^ 1][:len(nrecs)],
                                                                     done_event=self.free[slot ^ 1])
                self.pipe.reset()
                if multi:
                    vel = self.pipe.run_prefetched(handle)[0]
                else:
                    main.wait_event(self.ready[slot][0])
                    vel = self.pipe(views_of(slot, recs)[0], edges[0])[0].view(1, T, 3)
                    self.free[slot].record(main)
                out = self.h_vel[slot][: n * T].view(n, T, 3)
                out.copy_(vel, non_blocking=True)
            self.done[slot].record(main)
            if pending is not None:
                self.done[pending[0]].synchronize()
                yield pending[1]
            pending = (slot, out)
            cur, slot, handle = nxt, slot ^ 1, nxt_handle
        if pending is not None:
            self.done[pending[0]].synchronize()
            yield pending[1]


class StreamingSession:
    """Batch-1 streaming (BASELINE config 5, evfly_ros/run.py's 15-30 Hz loop): one window of events
    -> velocity command, with the whole device-side step (accumulate -> decode/crop -> percentile
    scale -> UNet/ConvLSTM -> ViT-LSTM, recurrent state carried) captured ONCE in a CUDA graph and
    replayed per window, so the ~200 kernel launches cost one graph launch.

    Event counts vary per window while a graph's kernel arguments are frozen, so the graph always
    scatters `capacity` records from a static buffer whose tail holds skip records (polarity 2)."""

    def __init__(self, pipe: "PerceptionPipeline", capacity: int = 131072, window_ns: int = 33_333_333, want_voxel: bool = True):
        self.pipe, self.cap = pipe, int(capacity)
        dev = pipe.dev
        skip = torch.zeros((self.cap, 16), dtype=torch.uint8, device=dev)
        skip[:, 12] = _lib.POL_SKIP
        self._skip = skip
        self.records = skip.clone()
        self._n_prev = 0
        self.edges = torch.tensor([0, window_ns], dtype=torch.int64, device=dev)
        self.want_voxel = want_voxel
        m = pipe.model
        self.h_unet = torch.zeros((1, 512, 8, 13), dtype=torch.float32, device=dev)
        self.c_unet = torch.zeros_like(self.h_unet)
        self.h_vit = torch.zeros((3, 128), dtype=torch.float32, device=dev)
        self.c_vit = torch.zeros_like(self.h_vit)
        self.desvel = torch.full((1, 1), pipe.desvel, dtype=torch.float32, device=dev)
        self.graph = None
        self._capture()

    def _step(self):
        frames, counts, voxel = self.pipe.frames_from_windows(self.records, self.edges, self.want_voxel, sorted_by_time=False)
        vel, (depth, _, ((hu, _), hv)) = self.pipe.model([frames, self.desvel, [[[self.h_unet, self.c_unet]], None], (self.h_vit, self.c_vit)])
        self.h_unet.copy_(hu[0][0]); self.c_unet.copy_(hu[0][1])
        self.h_vit.copy_(hv[0]); self.c_vit.copy_(hv[1])
        return vel, depth, counts, voxel

    def _capture(self):
        with torch.no_grad():
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):          # warm-up: packs weights, sets kernel attributes, sizes the allocator
                    self._step()
            torch.cuda.current_stream().wait_stream(s)
            self.reset()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.vel, self.depth, self.counts, self.voxel = self._step()
            self.reset()

    def reset(self):
        for t in (self.h_unet, self.c_unet, self.h_vit, self.c_vit):
            t.zero_()

    def load_events(self, records: torch.Tensor, t0_ns: int = 0, window_ns: int | None = None):
        """records uint8 [n,16] (device or pinned host); n <= capacity. Window = [t0, t0 + window)."""
        n = records.shape[0]
        if n > self.cap:
            raise _lib.EvflyError(f"window of {n} events exceeds the session capacity {self.cap}")
        self.records[:n].copy_(records, non_blocking=True)
        if n < self._n_prev:
            self.records[n:self._n_prev].copy_(self._skip[n:self._n_prev])
        self._n_prev = n
        if t0_ns != 0 or window_ns is not None:
            w = int(self.edges[1] - self.edges[0]) if window_ns is None else window_ns
            self.edges.copy_(torch.tensor([t0_ns, t0_ns + w], dtype=torch.int64), non_blocking=True)

    def step(self, records: torch.Tensor | None = None, **kw) -> torch.Tensor:
        """One window -> velocity command [1,3] (device tensor, valid until the next step)."""
        if records is not None:
            self.load_events(records, **kw)
        self.graph.replay()
        return self.vel


def build_deployed_model(device="cuda", seed_state_dict=None, logger=None):
    """OrigUNet_w_VITFLY_ViTLSTM in the shipped configuration (learner/configs/eval_config_real.txt:39-47:
    bev=2, skip_type=interp, num_recurrent=[1,0], resize_input=[260,346], velpred=0)."""
    from .learner_models import OrigUNet_w_VITFLY_ViTLSTM
    quiet = logger if logger is not None else (lambda *a, **k: None)
    enc = dict(num_layers=2, kernel_sizes=[5, 3], kernel_strides=[2, 2], out_channels=[8, 32], activations=["relu", "relu"],
               pool_type="max", invert_pool_inputs=True, pool_kernels=[2, 2], pool_strides=[2, 2], conv_function="conv2d")
    fc = dict(num_layers=4, layer_sizes=[1024, 128, 16, 1], activations=["leaky_relu"] * 3 + ["tanh"], dropout_p=0.1)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = OrigUNet_w_VITFLY_ViTLSTM(num_in_channels=2, num_out_channels=1, num_recurrent=[1, 0], input_shape=[1, 1, 260, 346],
                                      logger=quiet, velpred=0, enc_params=enc, fc_params=fc, form_BEV=2, evs_min_cutoff=1e-3,
                                      skip_type="interp", is_deployment=False)
    if seed_state_dict is not None:
        m.load_state_dict(seed_state_dict, strict=True)
    return m.to(device).eval().float()
