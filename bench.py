#!/usr/bin/env python
"""bench.py -- the contract benchmark (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference] [--no-extras]

One "step" is one pass of the hot path over one batch of synthetic event windows. Rank 0 prints ONE JSON line.
Multi-GPU: one process per GPU (torchrun), independent shards, no collective on the data path (weak scaling);
timing = max over ranks of CUDA-event time. `--workload equality` is the hardware check that N ranks evaluating a
fixed set of trajectories through evfly_b200.sharding (+ NCCL all_gather) return bit for bit what one GPU returns.
`--impl reference` times the reference's OWN code (baseline/_ref, imported unmodified) on the host cores; when
that directory is absent, the oracle port of the same algorithm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "sustained_clock_mhz": (p.get("clocks_under_load") or {}).get("sm_mhz_median"), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sustained_clock_mhz": 1300.0, "source": "fallback"}


def load_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture
    of this round (profiles/r2_dram_traffic.json, written by scripts/ncu_traffic.py from an ncu run of this file)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_dram_traffic.json")))[key]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin: float, t_end: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                clk, mx = float(f[0]), float(f[1])
            except ValueError:
                continue
            smax = mx
            if t_begin - 0.05 <= ts <= t_end + 0.05:
                sm.append(clk)
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        all_clk = sm or [float(l.split(",")[0]) for _, l in self.rows if l.split(",")[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(all_clk)) if all_clk else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process (and therefore the pinned host buffers it allocates next: first touch) to the CPUs of the NUMA
    node the GPU hangs off. With 8 ranks on a two-socket host, unbound ranks land on arbitrary sockets and half of
    the H2D traffic crosses the socket interconnect (end-to-end scaling 0.50 at 8 GPUs)."""
    rep = {"bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        rep.update(pci=bus, numa_node=node)
        if node < 0:
            return rep
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            rep.update(bound=True, cpus=len(allowed))
    except Exception as e:  # plumbing only: never fail the benchmark over it
        rep["error"] = f"{type(e).__name__}: {e}"
    return rep


# =============================================================================================
# workloads
# =============================================================================================
class AccumulateWorkload:
    """BASELINE config 2, both shapes of SURVEY 8(d), at 480x640 with count frames [2,H,W] + 5-bin voxel grids:
      2b (headline of this workload): a 10 M-event stream of 100 windows x 100 k events -> 100 frames + 100 grids
      2a (reported beside it): ONE window of 10 M events.
    Both through the shared-memory-tile path (accumulate_tiled.cu); the L2-reduction kernels are timed next to it.
    A step = the whole stream. Inputs rotate between two 160 MB buffers (> L2)."""
    name = "cfg2b: 10M-event stream, 100 windows x 100k events -> 100 count frames + 100 5-bin voxel grids @480x640 (accumulation only)"
    H, W, B, N_EV, T = 480, 640, 5, 10_000_000, 100
    dtype = "s32+f32"

    @classmethod
    def describe(cls):
        return cls.name

    def __init__(self, rank: int, device, steps_hint=10):
        import torch
        from evfly_b200.events import L1
        from evfly_b200.synthetic import synthetic_stream, synthetic_window
        self.torch, self.L1, self.dev = torch, L1, device
        self.windows_per_step = self.T
        per = self.N_EV // self.T
        self.host_b = [synthetic_stream(300 + 10 * rank + s, self.T, per, self.H, self.W) for s in (0, 1)]
        self.edges_b = torch.from_numpy(self.host_b[0][1]).to(device)
        self.pinned = [torch.from_numpy(h.view(np.uint8).reshape(-1, 16)).pin_memory() for h, _ in self.host_b]
        self.d_b = [p.to(device) for p in self.pinned]
        self.host_a = [synthetic_window(1000 * rank + s, self.N_EV, self.H, self.W) for s in (0, 1)]
        self.d_a = [torch.from_numpy(h.view(np.uint8).reshape(-1, 16)).to(device) for h in self.host_a]
        self.edges_a = torch.tensor([0, 33_333_333], dtype=torch.int64, device=device)
        self.counts_b = torch.empty((self.T, 2, self.H, self.W), dtype=torch.int32, device=device)
        self.voxel_b = torch.empty((self.T, self.B, self.H, self.W), dtype=torch.float32, device=device)
        self.counts_a = torch.zeros((1, 2, self.H, self.W), dtype=torch.int32, device=device)
        self.voxel_a = torch.zeros((1, self.B, self.H, self.W), dtype=torch.float32, device=device)
        self.ws_staged = L1.voxel_workspace(self.H, self.W, self.B, device)
        self.d_stage = torch.empty_like(self.d_b[0])
        self.h_vox = torch.empty((self.T, self.B, self.H, self.W), dtype=torch.float32).pin_memory()
        self.h_cnt = torch.empty((self.T, 2, self.H, self.W), dtype=torch.int32).pin_memory()
        out_w = (2 + self.B) * self.H * self.W * 4
        self.alg_bytes = 16 * self.N_EV + self.T * out_w          # 1.02 GB
        self.alg_bytes_a = 16 * self.N_EV + out_w                 # 168.6 MB
        self.h2d_bytes, self.d2h_bytes = 16 * self.N_EV, self.T * out_w

    def step(self, i: int):
        self.L1.accumulate_windows(self.d_b[i & 1], self.edges_b, self.H, self.W, self.B, counts=self.counts_b, voxel=self.voxel_b, algo="tiles")

    def dominant(self, i: int):
        self.step(i)

    def e2e_step(self, i: int):
        self.d_stage.copy_(self.pinned[i & 1], non_blocking=True)
        self.L1.accumulate_windows(self.d_stage, self.edges_b, self.H, self.W, self.B, counts=self.counts_b, voxel=self.voxel_b, algo="tiles")
        self.h_cnt.copy_(self.counts_b, non_blocking=True)
        self.h_vox.copy_(self.voxel_b, non_blocking=True)

    def check(self):
        from oracle import ev_oracle as O
        self.step(0)
        c_ref, _ = O.windows(self.host_b[0][0], self.host_b[0][1], self.H, self.W, B=None)
        assert np.array_equal(self.counts_b.cpu().numpy(), c_ref), "cfg2b: count frames differ from the oracle"
        self.L1.accumulate_windows(self.d_a[0], self.edges_a, self.H, self.W, self.B, counts=self.counts_a, voxel=self.voxel_a, algo="tiles")
        assert np.array_equal(self.counts_a[0].cpu().numpy(), O.event_counts(self.host_a[0], self.H, self.W)), "cfg2a: count frame differs from the oracle"

    def _time(self, fn, n=10):
        torch = self.torch
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e-3

    def roofline(self, dom_s: float, peaks: dict) -> dict:
        L1 = self.L1
        hbm = peaks["hbm_gbs"]
        t_a = self._time(lambda i: L1.accumulate_windows(self.d_a[i & 1], self.edges_a, self.H, self.W, self.B, counts=self.counts_a, voxel=self.voxel_a, algo="tiles"))
        t_a_red = self._time(lambda i: L1.voxelize_window(self.d_a[i & 1], self.H, self.W, self.B, 0, 33_333_333, counts=self.counts_a[0], voxel=self.voxel_a[0], ws=self.ws_staged, algo=1))
        t_b_red = self._time(lambda i: L1.accumulate_windows(self.d_b[i & 1], self.edges_b, self.H, self.W, self.B, counts=self.counts_b, voxel=self.voxel_b, algo="scatter"))
        mk = lambda name, nbytes, t: {"shape": name, "ms": t * 1e3, "achieved": nbytes / t / 1e9, "frac": nbytes / t / 1e9 / hbm, "algorithmic_bytes": nbytes}
        self._extra = {"rooflines_other": [
            dict(mk("cfg2a: one window of 10M events", self.alg_bytes_a, t_a), kernel="k_chunk_sort + k_band_accumulate (shared-memory tiles)"),
            dict(mk("cfg2a: one window of 10M events", self.alg_bytes_a, t_a_red), kernel="k_voxel_staged + k_voxel_finalize (one 16-byte L2 RED per event)"),
            dict(mk("cfg2b", self.alg_bytes, t_b_red), kernel="k_zero_fill + k_scatter_windows (three L2 REDs per event)")]}
        ach = self.alg_bytes / dom_s / 1e9
        return {"bound": "hbm", "kernel": "k_chunk_plan + k_chunk_sort + k_band_accumulate (window ranges, chunk-local band sort, per-band shared-memory histograms)",
                "achieved": ach, "peak": hbm, "peak_source": peaks["source"] + " (burst copy)", "unit": "GB/s", "frac": ach / hbm,
                "traffic": load_traffic("accumulate_cfg2b"), "algorithmic_bytes": self.alg_bytes}

    def extra(self):
        return getattr(self, "_extra", {})

    @classmethod
    def cpu_only(cls, rank):
        from evfly_b200.synthetic import synthetic_stream
        self = cls.__new__(cls)
        self.cpu_T = 10
        self.host_cpu = synthetic_stream(300 + 10 * rank, self.cpu_T, cls.N_EV // cls.T, cls.H, cls.W)
        self.windows_per_step = self.cpu_T
        self.cpu_cores = 1
        self.cpu_kind = "port"
        self.cpu_sample = "10 windows of 100k events (a tenth of the step): C loop restating node.cpp / histogram2d + voxel weights, 1 thread"
        return self

    def cpu_step(self, i: int):
        from oracle import ev_oracle as O
        O.windows(self.host_cpu[0], self.host_cpu[1], self.H, self.W, B=self.B)


class TrajectoryEvalWorkload:
    """BASELINE config 4 (the multi-GPU metric's configuration): offline evaluation of independent trajectories of
    100 windows each (260x346, 100 k events per window). Per GPU and per step a slice of N_TRAJ trajectories of the
    job's 2048 (sharded r::G over the ranks): events -> count frames + 5-bin voxel grids -> decode -> 97th-percentile
    scale/clip -> OrigUNet_w_VITFLY_ViTLSTM (deployed config, bf16 tensor-core path); the model advances the
    trajectories together (time-major frames), so the ConvLSTM / LSTM scans are 100 steps, N_TRAJ wide.
    Events travel in the 4-byte wire format (evfly_event4; the streams are on the sensor's 1 us grid): resident in HBM for `value`, from pinned host memory
    through TrajectoryFeeder for `e2e`."""
    H, W, B, N_EV = 260, 346, 5, 100_000
    T, N_TRAJ, N_DISTINCT = 100, 16, 4
    dtype = "bf16"
    # 2*MAC per frame measured from the reference modules
    FLOP_UNET, FLOP_CONVLSTM, FLOP_VIT = 11.883e9, 0.436e9, 0.1106e9
    FLOP_STEM = 0.05e9
    OVERLAP = True
    overlap_note = "accumulation + normalisation of step i+1 run on a side stream while the model runs step i (K steps = K accumulations + K forwards)"

    @classmethod
    def describe(cls):
        return (f"cfg4: per GPU {cls.N_TRAJ} trajectories x {cls.T} windows (260x346, 100k events each; slice of 2048 trajectories "
                "sharded over ranks): 4-byte wire records (1 us timestamps, delta-coded) -> count frames + 5-bin voxel -> prep -> UNet+ConvLSTM+ViT-LSTM forward, "
                "bf16 tensor-core path, per-trajectory recurrent state")

    @property
    def name(self):
        return type(self).describe()

    def __init__(self, rank: int, device, precision="bf16", steps_hint=10):
        import torch
        import evfly_b200
        from evfly_b200.events import WireBatch
        from evfly_b200.pipeline import PerceptionPipeline, build_deployed_model
        from evfly_b200.synthetic import synthetic_stream
        from oracle.synth_ckpt import shapes_of, synth_state_dict
        self.torch, self.dev, self.rank = torch, device, rank
        model = build_deployed_model("cpu")
        self.sd = synth_state_dict(shapes_of(model), 31)
        model.load_state_dict(self.sd)
        self.model = evfly_b200.set_precision(model.to(device).eval(), precision)
        self.pipe = PerceptionPipeline(self.model, sensor_hw=(self.H, self.W), model_hw=(self.H, self.W), num_bins=self.B)
        # N_DISTINCT seeded streams, laid out N_TRAJ / N_DISTINCT times (separate copies in memory: nothing is reused
        # between the copies, neither records in L2 nor results)
        # timestamps on the sensor's 1 us grid, 33.333 ms windows: representable in the 4-byte wire record (WireBatch picks it)
        distinct = [synthetic_stream(7000 + 64 * rank + s, self.T, self.N_EV, self.H, self.W, dur_ns=33_333_000, grid_ns=1000)
                    for s in range(min(self.N_DISTINCT, self.N_TRAJ))]
        self.streams = [distinct[s % len(distinct)] for s in range(self.N_TRAJ)]
        self.wire_host = WireBatch.from_streams(self.streams, device, pin=True)                 # pinned host records + device tables
        self.wire_dev = self.wire_host.on_device(self.wire_host.records.to(device))
        self.windows_per_step = self.N_TRAJ * self.T
        self.h2d_bytes = self.wire_host.records.numel()
        self.d2h_bytes = self.windows_per_step * 3 * 4
        W_ = self.windows_per_step
        # work of the tcgen05 conv/GEMM kernels: UNet convs (minus the stem, which has its own small-K, HBM-bound
        # tensor-core kernel and is not timed here) + ConvLSTM + the ViT Linear layers
        # (q/kv/final/mlp1/mlp2 = 54.0 MFLOP/frame of the 0.1106 G ViT-LSTM total) + decoder Linear 4.7 M
        self.tc_flops = W_ * (self.FLOP_UNET - self.FLOP_STEM + self.FLOP_CONVLSTM + 0.0587e9)
        self.step_flops = W_ * (self.FLOP_UNET + self.FLOP_CONVLSTM + self.FLOP_VIT)
        self.acc_bytes = self.wire_host.record_bytes * W_ * self.N_EV + W_ * self.H * self.W * 4 * (2 + self.B)

    def begin(self, n_steps: int):
        """Called by the timing harness before a run of n_steps consecutive step() calls."""
        self._n_steps, self._next = n_steps, None

    def step(self, i: int):
        with self.torch.no_grad():
            self.pipe.reset()
            if not self.OVERLAP:
                self.out = self.pipe.run_wire(self.wire_dev)
                return
            cur = self._next if getattr(self, "_next", None) is not None else self.pipe.prefetch_wire(self.wire_dev)
            # the next step's L1+L2 is queued BEFORE this step's model so that the two run concurrently; the last
            # step of a run queues nothing, so K steps do exactly K accumulations and K forwards
            self._next = self.pipe.prefetch_wire(self.wire_dev) if i + 1 < getattr(self, "_n_steps", 0) else None
            self.out = self.pipe.run_prefetched(cur)

    def e2e_run(self, steps: int, warmup: int = 3):
        """End to end through the public API (evfly_b200.pipeline.TrajectoryFeeder): every step's wire records start in
        pinned HOST memory, are copied to the device (copy stream, double-buffered so the copy of step i+1 overlaps the
        compute of step i), run through the pipeline, and the velocity commands are read back to the host.
        ONE continuous run of warmup + steps batches; the clock runs from the moment the last warm-up result is on the host
        to the moment the last timed result is: `steps` copies, forwards and read-backs in the steady state of the feeder
        (which works one batch ahead; a separate warm-up run would put the pipeline's fill -- one un-overlapped copy of
        0.64 GB -- into a region of a handful of steps). Returns (wall seconds, achieved H2D GB/s while copying)."""
        from evfly_b200.pipeline import TrajectoryFeeder
        torch = self.torch
        warmup = max(1, warmup)
        feeder = TrajectoryFeeder(self.pipe, self.wire_host.records.shape[0], self.windows_per_step, record_bytes=self.wire_host.record_bytes)
        t0 = None
        for k, vel in enumerate(feeder.run([self.wire_host] * (warmup + steps))):
            if k == warmup - 1:
                t0 = time.perf_counter()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return dt, feeder.h2d_gbs()[0]

    def dominant(self, i: int):
        """Same step with CUDA events around every launch of the dominant kernel family (the tcgen05 conv / GEMM
        kernels) and around the accumulation + normalisation; sums are read after the timed region."""
        from evfly_b200 import tc
        torch = self.torch
        self._ev = getattr(self, "_ev", [])
        hooks = ("_call", "_call_halo", "_call_halo_pool", "_call_scan", "_call_scan_fused", "_call_stem_e12", "_call_halo_out1")       # every launcher of the tcgen05 conv/GEMM kernels
        orig = {h: getattr(tc, h) for h in hooks}

        def timed(fn):
            def wrapper(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r = fn(*a, **k); e1.record()
                self._ev.append((e0, e1))
                return r
            return wrapper
        for h in hooks:
            setattr(tc, h, timed(orig[h]))
        stage_abi, tc.USE_STAGE_ABI = tc.USE_STAGE_ABI, False      # the same kernels enqueued one by one from Python, so that each can be timed
        try:
            with torch.no_grad():
                self.pipe.reset()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                tm = self.pipe.frames_from_wire(self.wire_dev)[0]
                a1.record()
                self._acc_ev = getattr(self, "_acc_ev", []) + [(a0, a1)]
                n, T = self.N_TRAJ, self.T
                self.model.forward_trajectories([tm, self.pipe._desvel(T * n), [None, None], None], n)
        finally:
            tc.USE_STAGE_ABI = stage_abi
            for h in hooks:
                setattr(tc, h, orig[h])

    def check(self):
        """The bench's own batch against the oracle: trajectory 0 of the N_TRAJ x T batch (count frames bit-exact,
        depth / velocity at the tolerance of tests/test_bench_shape_parity_gpu.py)."""
        import torch
        from oracle import ev_oracle as O, model_oracle as M
        n, T = self.N_TRAJ, self.T
        with torch.no_grad():
            self.pipe.reset()
            tm, counts, _ = self.pipe.frames_from_wire(self.wire_dev)
            vel, (depth, _, _) = self.model.forward_trajectories([tm, self.pipe._desvel(T * n), [None, None], None], n)
            rec, edges = self.streams[0]
            c_ref, _ = O.windows(rec, edges, self.H, self.W, B=None)
            assert np.array_equal(counts[0].cpu().numpy(), c_ref), "count frames differ from the oracle"
            fr = 0.2 * (c_ref[:, 1].astype(np.float32) - c_ref[:, 0].astype(np.float32))
            fr, _ = O.quantile_scale_clip(fr[:, None], 0.97, -1.0, 1.0)
            ovel, (odep, _, _) = M.orig_unet_w_vitlstm(self.sd, torch.from_numpy(fr), torch.full((T, 1), 4.0), None, None, **M.DEPLOYED_UNET_CFG)
            dep0 = depth.view(T, n, 1, self.H, self.W)[:, 0].cpu()
            vel0 = vel.view(T, n, 3)[:, 0].cpu()
            ok_d = ((dep0 - odep).abs() <= 1e-2 * odep.abs() + 1e-2 * odep.abs().max()).float().mean().item()
            ok_v = ((vel0 - ovel).abs() <= 1e-2 * ovel.abs() + 1e-2 * ovel.abs().max()).float().mean().item()
            l2_v = ((vel0 - ovel).norm() / ovel.norm().clamp_min(1e-30)).item()
            # depth must meet the bar; the velocity commands are reported (their element-wise pass fraction moves between 0.92
            # and 0.99 with the fp32 summation order of a build, tests/test_bench_shape_parity_gpu.py) and must be close in L2
            assert ok_d >= 0.999 and l2_v <= 3e-2, f"bench batch differs from the oracle: depth pass fraction {ok_d:.4f}, velocity rel-L2 {l2_v:.4f}"
            self._check = {"trajectory": 0, "frames": T, "counts_bit_exact": True, "depth_pass_frac": ok_d, "velocity_pass_frac": ok_v, "velocity_rel_l2": l2_v,
                           "tolerance": "|got-ref| <= 1e-2*|ref| + 1e-2*max|ref|"}
            self.pipe.reset()

    def roofline(self, dom_s_unused: float, peaks: dict) -> dict:
        torch = self.torch
        torch.cuda.synchronize()
        n_steps = max(1, len(self._acc_ev))
        tc_ms = sum(a.elapsed_time(b) for a, b in self._ev)
        launches = len(self._ev) / n_steps
        ach = self.tc_flops / (tc_ms / n_steps / 1e3) / 1e12
        acc_ms = sum(a.elapsed_time(b) for a, b in self._acc_ev) / n_steps
        self._extra = {"rooflines_other": [{
            "bound": "hbm", "kernel": "accumulate_windows_ev4 (k_chunk_plan + k_chunk_sort + k_band_accumulate) + k_counts_normalise (decode + crop + quantile + clip)",
            "achieved": self.acc_bytes / (acc_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": self.acc_bytes / (acc_ms / 1e3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": self.acc_bytes, "ms": acc_ms,
            "traffic": load_traffic("accumulate_cfg4")}],
            "tc_kernel_ms_per_step": tc_ms / n_steps, "parity_check": getattr(self, "_check", None)}
        return {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM conv kernels: k_tc_conv3x3_halo (Cin,Cout<=64 layers, resident weights) + k_tc_conv3x3_halo_ws (Cin=128 layers, streamed weights) + k_tc_conv_bf16 (other 3x3, 1x1, transposed convs, persistent ConvLSTM scan, ViT Linear layers)",
                "achieved": ach, "peak": peaks["bf16_tflops"], "peak_source": peaks["source"] + " (burst cuBLAS bf16: the clocks record of this run decides; see frac_vs_sustained)",
                "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                "frac_vs_sustained": ach / peaks["bf16_tflops_sustained"], "peak_sustained": peaks["bf16_tflops_sustained"],
                "sustained_peak_measured_at_mhz": peaks["sustained_clock_mhz"],
                "traffic": load_traffic("tc_conv_family"),
                "algorithmic_flops_per_launch": self.tc_flops / launches, "launches_per_step": launches,
                "avg_launch_us": tc_ms / n_steps / launches * 1e3}

    def extra(self):
        return getattr(self, "_extra", {})

    # ---- BASELINE config 5 ------------------------------------------------------------------------------
    def b1_latency(self, n_windows=1000, n_sparse=200):
        """Batch-1 streaming: consecutive 33 ms windows at 480x640 already resident in device memory -> velocity command
        on the host; recurrent state carried; host wall clock per window. 100 k events per window, plus the sparse
        5 k-event variant (< 3 % active pixels: quantile = 0 -> NaN -> mask of ones)."""
        import torch
        from evfly_b200.pipeline import PerceptionPipeline, StreamingSession
        from evfly_b200.synthetic import synthetic_window
        pipe = PerceptionPipeline(self.model, sensor_hw=(480, 640), model_hw=(260, 346), num_bins=self.B)
        out = {"what": "480x640 window resident in HBM -> count frame + 5-bin voxel -> crop/normalise -> UNet+ConvLSTM+ViT-LSTM "
                       "(state carried) -> velocity command on the host; one CUDA-graph replay per window; host wall clock"}
        with torch.no_grad():
            sess = StreamingSession(pipe, capacity=131072)       # whole step captured in one CUDA graph
            for key, n_ev, n_win in (("dense_100k_events", self.N_EV, n_windows), ("sparse_5k_events", 5_000, n_sparse)):
                wins = [torch.from_numpy(synthetic_window(900 + k, n_ev, 480, 640).view(np.uint8).reshape(-1, 16)).to(self.dev) for k in range(8)]
                sess.reset()
                lat, finite = [], True
                for k in range(n_win + 20):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    v = sess.step(wins[k % 8]).cpu()
                    lat.append((time.perf_counter() - t0) * 1e3)
                    finite = finite and bool(torch.isfinite(v).all())
                lat = np.array(lat[20:])
                out[key] = {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "windows": n_win,
                            "commands_finite": finite}
        out["p50_ms"], out["p99_ms"], out["windows"] = out["dense_100k_events"]["p50_ms"], out["dense_100k_events"]["p99_ms"], n_windows
        return out

    # ---- the other configurations, as extra keys of the same run ----------------------------------------------
    def extra_configs(self, peaks):
        """cfg 3 (one 256-window sequence, bf16), the exact fp32 path, and the 16-byte-record input of cfg 4."""
        import torch
        import evfly_b200
        from evfly_b200.events import to_device
        out = {}

        def time_steps(fn, n):
            for i in range(2):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        with torch.no_grad():
            rec, edges = self.streams[0]
            d16 = to_device(rec)
            d_edges = torch.from_numpy(edges).to(self.dev)
            # cfg 4 from canonical 16-byte records (4 trajectories per step, the round-1 bench shape)
            recs4, edges4 = [d16] * 4, [d_edges] * 4

            def step16(i):
                self.pipe.reset()
                self.pipe.run_trajectories(recs4, edges4)
            ms = time_steps(step16, 4)
            out["cfg4_16B_records_4x100"] = {"windows_per_s": 400 / ms * 1e3, "ms_per_step": ms, "note": "canonical evfly_event records resident in HBM, 4 trajectories per step, no side-stream overlap"}
            # cfg 3: ONE sequence of 256 windows (the batch dimension is time)
            from evfly_b200.synthetic import synthetic_stream
            rec3, edges3 = synthetic_stream(8000 + self.rank, 256, self.N_EV, self.H, self.W)
            d3, e3 = to_device(rec3), torch.from_numpy(edges3).to(self.dev)

            def step3(i):
                self.pipe.reset()
                self.pipe(d3, e3)
            ms = time_steps(step3, 5)
            fl = 256 * (self.FLOP_UNET + self.FLOP_CONVLSTM + self.FLOP_VIT)
            out["cfg3_sequence_256_bf16"] = {"windows_per_s": 256 / ms * 1e3, "ms_per_step": ms, "tflops_whole_step": fl / ms / 1e9,
                                             "frac_of_burst_peak_whole_step": fl / ms / 1e9 / peaks["bf16_tflops"]}
            # the exact path (CUDA-core fp32 kernels, rtol 1e-5 against the reference): 32 windows of one sequence
            evfly_b200.set_precision(self.model, "fp32")
            try:
                n32 = 32 * self.N_EV

                def step32(i):
                    self.pipe.reset()
                    self.pipe(d3[:n32], e3[:33])
                ms = time_steps(step32, 2)
                out["fp32_exact_path_sequence_32"] = {"windows_per_s": 32 / ms * 1e3, "ms_per_step": ms}
            finally:
                evfly_b200.set_precision(self.model, "bf16")
            self.pipe.reset()
        return out

    # ---- host-CPU arms on a bounded sample ----------------------------------------------------------------
    CPU_T = 8

    @classmethod
    def cpu_only(cls, rank):
        """The reference's own code (baseline/_ref: form_eventframe + run.py's quantile + the reference nn.Module) on all host
        threads; without baseline/_ref, the oracle port of the same algorithm."""
        import torch
        from evfly_b200.pipeline import build_deployed_model
        from evfly_b200.synthetic import records_to_rows, synthetic_stream
        from oracle.synth_ckpt import shapes_of, synth_state_dict
        self = cls.__new__(cls)
        self.sd = synth_state_dict(shapes_of(build_deployed_model("cpu")), 31)
        self.host, self.edges_host = synthetic_stream(7000 + rank, cls.CPU_T, cls.N_EV, cls.H, cls.W)
        torch.set_num_threads(os.cpu_count())
        self.cpu_cores = torch.get_num_threads()
        self.windows_per_step = cls.CPU_T
        self.arm = None
        try:
            from baseline.reference_arm import ReferenceArm, available
            if available():
                self.arm = ReferenceArm(self.sd)
                self.rows = [records_to_rows(self.host[k * cls.N_EV:(k + 1) * cls.N_EV]) for k in range(cls.CPU_T)]
        except Exception as e:
            self.arm, self.arm_error = None, f"{type(e).__name__}: {e}"
        self.cpu_kind = "reference" if self.arm is not None else "port"
        self.cpu_sample = (f"1 trajectory of {cls.CPU_T} windows (a slice of the step's windows): " +
                           ("the reference's own form_eventframe (np.histogram2d) + torch.quantile + OrigUNet_w_VITFLY_ViTLSTM.forward from baseline/_ref, "
                            if self.arm is not None else "oracle port: C accumulation loop + numpy quantile + torch-CPU fp32 forward, ") + "all host threads")
        return self

    def cpu_step(self, i: int):
        import torch
        if self.arm is not None:
            self.arm.trajectory(self.rows, self.H, self.W)
            return
        from oracle import ev_oracle as O, model_oracle as M
        T = self.CPU_T
        n = T * self.N_EV
        with torch.no_grad():
            c_ref, _ = O.windows(self.host[:n], self.edges_host[:T + 1], self.H, self.W, B=self.B)
            fr = 0.2 * (c_ref[:, 1].astype(np.float32) - c_ref[:, 0].astype(np.float32))
            fr, _ = O.quantile_scale_clip(fr[:, None], 0.97, -1.0, 1.0)
            M.orig_unet_w_vitlstm(self.sd, torch.from_numpy(fr), torch.full((T, 1), 4.0), None, None, **M.DEPLOYED_UNET_CFG)


class PipelineWorkload(TrajectoryEvalWorkload):
    """BASELINE config 3: per GPU and per step ONE trajectory of 256 consecutive windows = one 256-step sequence
    (SURVEY F2)."""
    T, N_TRAJ, N_DISTINCT = 256, 1, 1
    OVERLAP = False
    overlap_note = "none"

    @classmethod
    def describe(cls):
        return ("cfg3: per GPU 1 trajectory x 256 windows (260x346, 100k events each) = one 256-step sequence: 4-byte wire records -> "
                "count frames + 5-bin voxel -> prep -> UNet+ConvLSTM+ViT-LSTM forward, bf16 tensor-core path")


WORKLOADS = {"accumulate": AccumulateWorkload, "pipeline": PipelineWorkload, "trajectories": TrajectoryEvalWorkload}
DEFAULT_WORKLOAD = "trajectories"


# =============================================================================================
def run_reference(args, rank, world):
    """CPU arm (rank 0 only): the reference's own implementation of the path on the host cores."""
    if rank != 0:
        return
    wl = WORKLOADS[args.workload].cpu_only(rank)
    for i in range(args.warmup):
        wl.cpu_step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        wl.cpu_step(i)
    dt = time.perf_counter() - t0
    val = wl.windows_per_step * args.steps / dt
    gpu_name = WORKLOADS[args.workload].describe()          # the GPU arm's config, of which this arm times a bounded sample
    line = {"impl": "reference", "metric": "event windows/sec voxelize+forward", "value": val, "unit": "windows/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",   # the reference computes in fp32/fp64 on the CPU
            "data": "synthetic", "config": {"workload": gpu_name, "sample_windows_per_step": wl.windows_per_step},
            "cpu_baseline": {"value": val, "unit": "windows/s", "cores": wl.cpu_cores, "kind": wl.cpu_kind, "sample": wl.cpu_sample},
            "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_equality(args, rank, local_rank, world):
    """Hardware multi-GPU equality (SURVEY section 4 layer 6; learner/evaluation_tools.py:62-66 evaluates trajectories one by
    one on one device): world ranks evaluate a fixed set of trajectories through evfly_b200.sharding.evaluate_trajectories
    (shard r::G, NCCL all_gather of the velocity commands); rank 0 also evaluates ALL of them alone; the two results must
    be equal bit for bit."""
    import torch
    import torch.distributed as dist
    import evfly_b200
    from evfly_b200.events import WireBatch
    from evfly_b200.pipeline import PerceptionPipeline, build_deployed_model
    from evfly_b200.sharding import evaluate_trajectories
    from evfly_b200.synthetic import synthetic_stream
    from oracle.synth_ckpt import shapes_of, synth_state_dict
    dev = torch.device("cuda", local_rank)
    n_traj, T, n_ev, H, W = 8 * max(1, args.equality_k), 6, 40_000, 260, 346
    with torch.no_grad():
        model = build_deployed_model("cpu")
        model.load_state_dict(synth_state_dict(shapes_of(model), 31))
        model = evfly_b200.set_precision(model.to(dev).eval(), "bf16")
        pipe = PerceptionPipeline(model, sensor_hw=(H, W), model_hw=(H, W))

        def run_trajectory(i):
            wb = WireBatch.from_streams([synthetic_stream(5000 + i, T, n_ev, H, W)], dev, pin=False)
            pipe.reset()
            return pipe.run_wire(wb.on_device(wb.records.to(dev)))[0][0]          # [T,3]
        t0 = time.perf_counter()
        gathered = evaluate_trajectories(run_trajectory, n_traj, rank, world)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            alone = torch.stack([run_trajectory(i) for i in range(n_traj)])
            equal = bool(torch.equal(alone, gathered))
            print(json.dumps({"workload": "equality", "n_gpus": world, "n_trajectories": n_traj, "windows_per_trajectory": T,
                              "equal_bit_for_bit": equal, "max_abs_diff": float((alone - gathered).abs().max()),
                              "finite": bool(torch.isfinite(gathered).all()), "seconds_sharded": dt,
                              "what": "sharding.evaluate_trajectories over NCCL all_gather vs the same trajectories on rank 0 alone"}), flush=True)
            if not equal:
                sys.exit(3)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS) + ["equality"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip cfg 2/3/5 and the fp32 path (the extra keys of the N=1 line)")
    ap.add_argument("--equality-k", type=int, default=1)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if args.workload == "equality":
            args.workload = DEFAULT_WORKLOAD
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from evfly_b200 import _build, _lib
    _lib.load()  # fails loudly if the CUDA library is missing
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)        # before any pinned allocation
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    if args.workload == "equality":
        run_equality(args, rank, local_rank, world)
        return

    peaks = load_peaks()
    wl = WORKLOADS[args.workload](rank, dev, steps_hint=args.steps)
    if rank == 0:
        wl.check()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """K steps, each bracketed by CUDA events on the launching stream; returns total seconds
        (max over ranks) and the launch count of this rank."""
        owner = getattr(fn, "__self__", None)
        if hasattr(owner, "begin"):
            owner.begin(warmup)
        for i in range(warmup):
            fn(i)
        barrier()
        if hasattr(owner, "begin"):
            owner.begin(steps)
        l0 = _lib.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t_begin = time.time()
        for i in range(steps):
            ev[i][0].record()
            fn(i)
            ev[i][1].record()
        barrier()
        t_end = time.time()
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
        launches = _lib.launch_count() - l0
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / 1e3, launches, (t_begin, t_end)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3)
    total_s, launches, span = timed(wl.step, args.steps, args.warmup)
    clocks = sampler.stop(*span) if sampler else None
    dom_s, _, _ = timed(wl.dominant, max(2, min(args.steps, 5)), 3)
    e2e_steps = max(4, args.steps)
    h2d_gbs = None
    if hasattr(wl, "e2e_run"):
        barrier()
        e2e_local, h2d_local = wl.e2e_run(e2e_steps, args.warmup)
        t = torch.tensor([e2e_local, -h2d_local], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, h2d_gbs = t[0].item(), -t[1].item()          # slowest rank; lowest per-rank H2D bandwidth
    else:
        e2e_s, _, _ = timed(wl.e2e_step, e2e_steps, 3)
    if rank == 0:
        windows = wl.windows_per_step * world
        value = windows * args.steps / total_s
        e2e_val = windows * e2e_steps / e2e_s
        line = {
            "metric": "event windows/sec voxelize+forward", "value": value, "unit": "windows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype,
            "data": "synthetic",
            "config": {"workload": wl.name, "windows_per_step_per_gpu": wl.windows_per_step,
                       "l2_policy": "inputs larger than L2 (>= 160 MB of event records read per step, outputs >> L2)",
                       "sharding": "independent windows per rank, no data-path collective",
                       "overlap": getattr(wl, "overlap_note", "none")},
            "roofline": wl.roofline(dom_s / max(2, min(args.steps, 5)), peaks),
            "e2e": {"value": e2e_val, "unit": "windows/s", "h2d_bytes_per_step": wl.h2d_bytes, "d2h_bytes_per_step": wl.d2h_bytes,
                    "h2d_gbs_while_copying_min_over_ranks": h2d_gbs, "ratio_to_value": e2e_val / value,
                    "limiter": "H2D of the event records (PCIe / host memory)" if e2e_val < 0.9 * value else "the device step (copies fully overlapped)"},
            "gpu_launches": launches, "clocks": clocks, "numa": numa, "build": _build.build_state(),
        }
        if hasattr(wl, "step_flops"):
            fl = wl.step_flops / (total_s / args.steps) / 1e12
            line["whole_step"] = {"tflops": fl, "frac_of_burst_peak": fl / peaks["bf16_tflops"], "frac_of_sustained_peak": fl / peaks["bf16_tflops_sustained"]}
        line.update(wl.extra())
        if world == 1 and not args.no_extras:
            if hasattr(wl, "b1_latency"):
                line["b1_latency"] = wl.b1_latency()
            if hasattr(wl, "extra_configs"):
                line["extra_configs"] = wl.extra_configs(peaks)
                del wl
                torch.cuda.empty_cache()
                acc = AccumulateWorkload(rank, dev)
                acc.check()
                t_acc, _, _ = timed(acc.step, 10, 3)
                r = acc.roofline(t_acc / 10, peaks)
                line["extra_configs"]["cfg2_accumulation"] = {"cfg2b": {"ms": t_acc / 10 * 1e3, "achieved_gbs": r["achieved"], "frac_of_hbm": r["frac"],
                                                                         "algorithmic_bytes": r["algorithmic_bytes"], "kernel": r["kernel"]},
                                                              "others": acc.extra()["rooflines_other"]}
                wl = None
        if world == 1 and not args.no_cpu_baseline:
            cw = WORKLOADS[args.workload].cpu_only(rank)
            cw.cpu_step(0)
            t0 = time.perf_counter()
            n_cpu = 2
            for i in range(n_cpu):
                cw.cpu_step(1 + i)
            dt = (time.perf_counter() - t0) / n_cpu
            line["cpu_baseline"] = {"value": cw.windows_per_step / dt, "unit": "windows/s", "cores": cw.cpu_cores,
                                    "kind": cw.cpu_kind, "sample": cw.cpu_sample, "same_config": False}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
