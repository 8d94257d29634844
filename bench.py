#!/usr/bin/env python
"""bench.py -- the contract benchmark (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" is one pass of the hot path over one batch of synthetic event windows. Rank 0
prints ONE JSON line. Multi-GPU: one process per GPU (torchrun), independent shards, no
collective on the data path (weak scaling); timing = max over ranks of CUDA-event time.
`--impl reference` times the CPU port of the reference's algorithm (oracle/) on host cores.
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
                "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


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


# =============================================================================================
# workloads
# =============================================================================================
class AccumulateWorkload:
    """BASELINE config 2a: one window of 10 M events at 480x640 -> int32 count frames [2,H,W] +
    5-bin fp32 voxel grid [5,H,W]. A step = one window. Inputs rotate between two 160 MB
    buffers (> L2) so no step finds its events in cache."""
    name = "cfg2a: 10M-event window -> count frames + 5-bin voxel grid @480x640 (accumulation only)"
    H, W, B, N_EV = 480, 640, 5, 10_000_000
    T0, T1 = 0, 33_333_333
    windows_per_step = 1
    dtype = "s32+f32"

    def __init__(self, rank: int, device):
        import torch
        from evfly_b200.events import L1
        from evfly_b200.synthetic import synthetic_window
        self.torch, self.L1, self.dev = torch, L1, device
        self.host = [synthetic_window(1000 * rank + s, self.N_EV, self.H, self.W) for s in (0, 1)]
        self.pinned = [torch.from_numpy(h.view(np.uint8).reshape(-1, 16)).pin_memory() for h in self.host]
        self.d_in = [p.to(device) for p in self.pinned]
        self.counts = torch.zeros((2, self.H, self.W), dtype=torch.int32, device=device)
        self.voxel = torch.zeros((self.B, self.H, self.W), dtype=torch.float32, device=device)
        self.ws = L1.voxel_workspace(self.H, self.W, self.B, device)
        self.d_stage = torch.empty_like(self.d_in[0])
        self.h_counts = torch.empty((2, self.H, self.W), dtype=torch.int32).pin_memory()
        self.h_voxel = torch.empty((self.B, self.H, self.W), dtype=torch.float32).pin_memory()
        self.alg_bytes = 16 * self.N_EV + (2 + self.B) * self.H * self.W * 4
        self.h2d_bytes = 16 * self.N_EV
        self.d2h_bytes = (2 + self.B) * self.H * self.W * 4

    def step(self, i: int):
        self.L1.voxelize_window(self.d_in[i & 1], self.H, self.W, self.B, self.T0, self.T1,
                                counts=self.counts, voxel=self.voxel, ws=self.ws, algo=1)

    # the dominant kernel is the whole step here (scatter + finalise are timed together)
    def dominant(self, i: int):
        self.step(i)

    def e2e_step(self, i: int):
        self.d_stage.copy_(self.pinned[i & 1], non_blocking=True)
        self.L1.voxelize_window(self.d_stage, self.H, self.W, self.B, self.T0, self.T1,
                                counts=self.counts, voxel=self.voxel, ws=self.ws, algo=1)
        self.h_counts.copy_(self.counts, non_blocking=True)
        self.h_voxel.copy_(self.voxel, non_blocking=True)

    def check(self):
        from oracle import ev_oracle as O
        self.step(0)
        c_ref = O.event_counts(self.host[0], self.H, self.W)
        assert np.array_equal(self.counts.cpu().numpy(), c_ref), "bench output differs from the oracle"

    def roofline(self, dom_s: float, peaks: dict) -> dict:
        ach = self.alg_bytes / dom_s / 1e9
        return {"bound": "hbm", "kernel": "k_voxel_staged + k_voxel_finalize", "achieved": ach,
                "peak": peaks["hbm_gbs"], "peak_source": peaks["source"] + " (burst copy)", "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes": self.alg_bytes}

    @classmethod
    def cpu_only(cls, rank):
        from evfly_b200.synthetic import synthetic_window
        self = cls.__new__(cls)
        self.host = [synthetic_window(1000 * rank + s, cls.N_EV, cls.H, cls.W) for s in (0, 1)]
        return self

    def extra(self):
        return {}

    # CPU port of the reference algorithm (oracle): one full window, single thread
    def cpu_step(self, i: int):
        from oracle import ev_oracle as O
        O.voxel_window(self.host[i & 1], self.H, self.W, self.B, self.T0, self.T1)
    cpu_sample = "1 window of 10M events (the full step), C loop restating node.cpp / histogram2d, 1 thread"
    cpu_cores = 1


class PipelineWorkload:
    """voxelize + forward (BASELINE metric): per GPU and per step ONE trajectory of 256 consecutive
    33 ms event windows (cfg-1-style: 260x346, 100k events each = 25.6 M events, 410 MB of records)
    -> int32 count frames + 5-bin voxel grids (accumulate_windows) -> decode -> 97th-percentile
    scale/clip -> OrigUNet_w_VITFLY_ViTLSTM (deployed config, bf16 tensor-core path) over the
    256-step sequence with fresh recurrent state. Configs 3/4 of BASELINE.json: the 256 frames are
    one sequence; ranks process independent trajectories (weak scaling)."""
    name = ("cfg3/4: per GPU 1 trajectory x 256 windows (260x346, 100k events each): count frames + 5-bin voxel "
            "-> prep -> UNet+ConvLSTM+ViT-LSTM forward, bf16 tensor-core path")
    H, W, B, T, N_EV = 260, 346, 5, 256, 100_000
    N_TRAJ = 1
    windows_per_step = 256
    dtype = "bf16"
    # 2*MAC per frame measured from the reference modules
    FLOP_UNET, FLOP_CONVLSTM, FLOP_VIT = 11.883e9, 0.436e9, 0.1106e9
    FLOP_STEM, FLOP_OUT = 0.05e9, 0.0026e9

    def __init__(self, rank: int, device, precision="bf16"):
        import torch
        import evfly_b200
        from evfly_b200.pipeline import PerceptionPipeline, build_deployed_model
        from evfly_b200.synthetic import synthetic_stream
        from oracle.synth_ckpt import shapes_of, synth_state_dict
        self.torch, self.dev = torch, device
        model = build_deployed_model("cpu")
        self.sd = synth_state_dict(shapes_of(model), 31)
        model.load_state_dict(self.sd)
        self.model = evfly_b200.set_precision(model.to(device).eval(), precision)
        self.pipe = PerceptionPipeline(self.model, sensor_hw=(self.H, self.W), model_hw=(self.H, self.W), num_bins=self.B)
        streams = [synthetic_stream(7000 + 64 * rank + s, self.T, self.N_EV, self.H, self.W) for s in range(self.N_TRAJ)]
        self.host, self.edges_host = streams[0]
        self.pinned_list = [torch.from_numpy(h.view(np.uint8).reshape(-1, 16)).pin_memory() for h, _ in streams]
        self.d_in_list = [p.to(device) for p in self.pinned_list]
        self.d_edges = torch.from_numpy(self.edges_host).to(device)
        self.d_in = self.d_in_list[0]
        self.h2d_bytes = sum(p.numel() for p in self.pinned_list)
        self.d2h_bytes = self.N_TRAJ * self.T * 3 * 4
        self.windows_per_step = self.N_TRAJ * self.T
        # work of the tcgen05 conv/GEMM kernels: UNet convs (minus the stem, which has its own small-K, HBM-bound
        # tensor-core kernel and is not timed here) + ConvLSTM + the ViT Linear layers
        # (q/kv/final/mlp1/mlp2 = 54.0 MFLOP/frame of the 0.1106 G ViT-LSTM total) + decoder Linear 4.7 M
        W_ = self.N_TRAJ * self.T
        self.tc_flops = W_ * (self.FLOP_UNET - self.FLOP_STEM + self.FLOP_CONVLSTM + 0.0587e9)
        self.acc_bytes = 16 * W_ * self.N_EV + W_ * self.H * self.W * 4 * (2 + self.B)

    overlap_note = "cfg4: accumulation + normalisation of step i+1 run on a side stream while the model runs step i (K steps = K accumulations + K forwards)"
    OVERLAP = True      # cfg 4: accumulate + normalise batch i+1 on a side stream while the model runs batch i

    def begin(self, n_steps: int):
        """Called by the timing harness before a run of n_steps consecutive step() calls."""
        self._n_steps, self._next = n_steps, None

    def step(self, i: int):
        with self.torch.no_grad():
            self.pipe.reset()
            if self.N_TRAJ == 1:
                self.out = self.pipe(self.d_in, self.d_edges)
            elif not self.OVERLAP:
                self.out = self.pipe.run_trajectories(self.d_in_list, [self.d_edges] * self.N_TRAJ)
            else:
                edges = [self.d_edges] * self.N_TRAJ
                cur = self._next if getattr(self, "_next", None) is not None else self.pipe.prefetch_trajectories(self.d_in_list, edges)
                # the next step's L1+L2 is queued BEFORE this step's model so that the two run concurrently; the last
                # step of a run queues nothing, so K steps do exactly K accumulations and K forwards
                self._next = self.pipe.prefetch_trajectories(self.d_in_list, edges) if i + 1 < getattr(self, "_n_steps", 0) else None
                self.out = self.pipe.run_prefetched(cur)

    def e2e_run(self, steps: int):
        """End to end through the public API (evfly_b200.pipeline.TrajectoryFeeder): every step's 410 MB of
        records start in pinned HOST memory, are copied to the device (copy stream, double-buffered so the
        copy of step i+1 overlaps the compute of step i), run through the pipeline, and the velocity commands
        are read back to the host. Returns wall seconds for `steps` steps (first copy included)."""
        from evfly_b200.pipeline import TrajectoryFeeder
        torch = self.torch
        feeder = TrajectoryFeeder(self.pipe, sum(p.shape[0] for p in self.pinned_list), self.N_TRAJ * self.T)
        one = (self.pinned_list, [self.d_edges] * self.N_TRAJ) if self.N_TRAJ > 1 else (self.pinned_list[0], self.d_edges)
        batches = [one] * steps
        for _ in feeder.run([one] * 2):      # warm-up
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for vel in feeder.run(batches):
            pass
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def dominant(self, i: int):
        """Same step with CUDA events around every launch of the dominant kernel (k_tc_conv_bf16)
        and around the accumulation call; sums are read after the timed region."""
        from evfly_b200 import tc
        torch = self.torch
        self._ev = getattr(self, "_ev", [])
        hooks = ("_call", "_call_halo", "_call_halo_pool", "_call_scan")       # every launcher of the tcgen05 conv/GEMM kernels
        orig = {h: getattr(tc, h) for h in hooks}

        def timed(fn):
            def wrapper(*a):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(*a); e1.record()
                self._ev.append((e0, e1))
            return wrapper
        for h in hooks:
            setattr(tc, h, timed(orig[h]))
        try:
            with torch.no_grad():
                self.pipe.reset()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                if self.N_TRAJ == 1:
                    fr = self.pipe.frames_from_windows(self.d_in_list[0], self.d_edges)[0]
                else:
                    tm = self.pipe.frames_from_trajectories(self.d_in_list, [self.d_edges] * self.N_TRAJ)[0]
                a1.record()
                self._acc_ev = getattr(self, "_acc_ev", []) + [(a0, a1)]
                if self.N_TRAJ == 1:
                    self.pipe.forward(fr)
                else:
                    n, T = self.N_TRAJ, self.T
                    dv = torch.full((T * n, 1), 4.0, dtype=torch.float32, device=self.dev)
                    self.model.forward_trajectories([tm, dv, [None, None], None], n)
        finally:
            for h in hooks:
                setattr(tc, h, orig[h])

    def check(self):
        """one short sequence against the oracle (bf16 tolerance, tests/test_models_bf16_gpu.py)"""
        import torch
        from oracle import ev_oracle as O, model_oracle as M
        T = 4
        n = T * self.N_EV
        with torch.no_grad():
            self.pipe.reset()
            vel, depth, counts, voxel = self.pipe(self.d_in[:n], self.d_edges[:T + 1])
            c_ref, _ = O.windows(self.host[:n], self.edges_host[:T + 1], self.H, self.W, B=None)
            assert np.array_equal(counts.cpu().numpy(), c_ref), "count frames differ from the oracle"
            fr = 0.2 * (c_ref[:, 1].astype(np.float32) - c_ref[:, 0].astype(np.float32))
            fr, _ = O.quantile_scale_clip(fr[:, None], 0.97, -1.0, 1.0)
            ovel, (odep, _, _) = M.orig_unet_w_vitlstm(self.sd, torch.from_numpy(fr), torch.full((T, 1), 4.0), None, None, **M.DEPLOYED_UNET_CFG)
            l2 = float(torch.linalg.norm(depth.cpu() - odep) / torch.linalg.norm(odep))
            assert l2 < 1e-2, f"depth differs from the oracle: rel L2 {l2}"
            assert float((vel.cpu() - ovel).abs().max()) < 1e-2 * float(ovel.abs().mean() + ovel.abs().max())
            self.pipe.reset()

    def roofline(self, dom_s_unused: float, peaks: dict) -> dict:
        torch = self.torch
        torch.cuda.synchronize()
        n_steps = max(1, len(self._acc_ev))
        tc_ms = sum(a.elapsed_time(b) for a, b in self._ev)
        from evfly_b200 import tc
        # the ConvLSTM scan is one timed call: one persistent launch, or T step launches
        launches = len(self._ev) / n_steps + (0 if tc.PERSISTENT_SCAN else self.T - 1)
        ach = self.tc_flops / (tc_ms / n_steps / 1e3) / 1e12
        acc_ms = sum(a.elapsed_time(b) for a, b in self._acc_ev) / n_steps
        self._extra = {"rooflines_other": [{
            "bound": "hbm", "kernel": "accumulate_windows (k_zero_fill + k_scatter_windows) + decode + quantile",
            "achieved": self.acc_bytes / (acc_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": self.acc_bytes / (acc_ms / 1e3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": self.acc_bytes, "ms": acc_ms}],
            "tc_kernel_ms_per_step": tc_ms / n_steps}
        return {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM conv kernels: k_tc_conv3x3_halo (Cin,Cout<=64 layers, resident weights) + k_tc_conv3x3_halo_ws (Cin=128 layers, streamed weights) + k_tc_conv_bf16 (other 3x3, 1x1, transposed convs, persistent ConvLSTM scan, ViT Linear layers)",
                "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "peak_source": peaks["source"] + " (sustained cuBLAS bf16; kernel timed inside a long step)",
                "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"], "traffic": None,
                "algorithmic_flops_per_launch": self.tc_flops / launches, "launches_per_step": launches,
                "avg_launch_us": tc_ms / n_steps / launches * 1e3}

    def extra(self):
        return getattr(self, "_extra", {})

    def b1_latency(self, n_windows=200):
        """BASELINE config 5: batch-1 streaming, 33 ms windows of 100k events at 480x640 already resident
        in device memory -> velocity command on the host; recurrent state carried; wall clock."""
        import torch
        from evfly_b200.pipeline import PerceptionPipeline
        from evfly_b200.synthetic import synthetic_window
        from evfly_b200.pipeline import StreamingSession
        pipe = PerceptionPipeline(self.model, sensor_hw=(480, 640), model_hw=(260, 346), num_bins=self.B)
        wins = [torch.from_numpy(synthetic_window(900 + k, self.N_EV, 480, 640).view(np.uint8).reshape(-1, 16)).to(self.dev) for k in range(8)]
        with torch.no_grad():
            sess = StreamingSession(pipe, capacity=131072)       # whole step captured in one CUDA graph
            lat = []
            for k in range(n_windows + 20):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                v = sess.step(wins[k % 8]).cpu()
                lat.append((time.perf_counter() - t0) * 1e3)
        lat = np.array(lat[20:])
        return {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "windows": n_windows,
                "what": "480x640 window of 100k events resident in HBM -> count frame + 5-bin voxel -> crop/normalise -> "
                        "UNet+ConvLSTM+ViT-LSTM (state carried) -> velocity command on the host; one CUDA-graph replay per window"}

    # ---- CPU port of the reference algorithm (oracle/) on a bounded sample --------------------------
    CPU_T = 8
    cpu_sample = "1 trajectory of 8 windows (a slice of the step's windows): C accumulation loop + numpy quantile + torch-CPU fp32 forward, all host threads"

    @classmethod
    def cpu_only(cls, rank):
        import torch
        from evfly_b200.pipeline import build_deployed_model
        from evfly_b200.synthetic import synthetic_stream
        from oracle.synth_ckpt import shapes_of, synth_state_dict
        self = cls.__new__(cls)
        self.sd = synth_state_dict(shapes_of(build_deployed_model("cpu")), 31)
        self.host, self.edges_host = synthetic_stream(7000 + rank, cls.CPU_T, cls.N_EV, cls.H, cls.W)
        torch.set_num_threads(os.cpu_count())
        self.cpu_cores = torch.get_num_threads()
        self.windows_per_step = cls.CPU_T
        return self

    def cpu_step(self, i: int):
        import torch
        from oracle import ev_oracle as O, model_oracle as M
        T = self.CPU_T
        n = T * self.N_EV
        with torch.no_grad():
            c_ref, _ = O.windows(self.host[:n], self.edges_host[:T + 1], self.H, self.W, B=self.B)
            fr = 0.2 * (c_ref[:, 1].astype(np.float32) - c_ref[:, 0].astype(np.float32))
            fr, _ = O.quantile_scale_clip(fr[:, None], 0.97, -1.0, 1.0)
            M.orig_unet_w_vitlstm(self.sd, torch.from_numpy(fr), torch.full((T, 1), 4.0), None, None, **M.DEPLOYED_UNET_CFG)


class TrajectoryEvalWorkload(PipelineWorkload):
    """BASELINE config 4 (the multi-GPU metric's configuration): offline evaluation of independent trajectories of
    100 windows each. Per GPU and per step, a slice of 4 trajectories (of the job's 2048, sharded r::G over the
    ranks) = 400 windows: each trajectory's stream -> count frames + voxel grids -> normalise; the model then
    advances the 4 trajectories together (time-major frames), so the ConvLSTM/LSTM scans are 100 steps wide-4."""
    name = ("cfg4: per GPU 4 trajectories x 100 windows (260x346, 100k events each; slice of 2048 trajectories sharded over "
            "ranks): count frames + 5-bin voxel -> prep -> UNet+ConvLSTM+ViT-LSTM forward, bf16 tensor-core path, "
            "per-trajectory recurrent state")
    T, N_TRAJ = 100, 4


WORKLOADS = {"accumulate": AccumulateWorkload, "pipeline": PipelineWorkload, "trajectories": TrajectoryEvalWorkload}
DEFAULT_WORKLOAD = "trajectories"


# =============================================================================================
def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's algorithm on the host cores (rank 0 only)."""
    if rank != 0:
        return
    wl_cls = WORKLOADS[args.workload]
    wl = wl_cls.cpu_only(rank)
    for i in range(args.warmup):
        wl.cpu_step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        wl.cpu_step(i)
    dt = time.perf_counter() - t0
    val = wl.windows_per_step * args.steps / dt
    line = {"impl": "reference", "metric": "event windows/sec voxelize+forward", "value": val, "unit": "windows/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",   # the reference computes in fp32/fp64 on the CPU
            "data": "synthetic", "config": {"workload": wl.name},
            "cpu_baseline": {"value": val, "unit": "windows/s", "cores": wl.cpu_cores, "kind": "port", "sample": wl.cpu_sample},
            "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from evfly_b200 import _lib
    _lib.load()  # fails loudly if the CUDA library is missing
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    peaks = load_peaks()
    wl = WORKLOADS[args.workload](rank, dev)
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
    dom_s, _, _ = timed(wl.dominant, args.steps, args.warmup)
    e2e_steps = max(4, args.steps)
    if hasattr(wl, "e2e_run"):
        barrier()
        e2e_local = wl.e2e_run(e2e_steps)
        t = torch.tensor([e2e_local], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    else:
        e2e_s, _, _ = timed(wl.e2e_step, e2e_steps, 3)
    if rank == 0:
        windows = wl.windows_per_step * world
        value = windows * args.steps / total_s
        line = {
            "metric": "event windows/sec voxelize+forward", "value": value, "unit": "windows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype,
            "data": "synthetic",
            "config": {"workload": wl.name, "windows_per_step_per_gpu": wl.windows_per_step,
                       "l2_policy": "inputs larger than L2 (>= 160 MB of event records read per step, outputs >> L2)",
                       "sharding": "independent windows per rank, no data-path collective",
                       "overlap": getattr(wl, "overlap_note", "none")},
            "roofline": wl.roofline(dom_s / args.steps, peaks),
            "e2e": {"value": windows * e2e_steps / e2e_s, "unit": "windows/s",
                    "h2d_bytes_per_step": wl.h2d_bytes, "d2h_bytes_per_step": wl.d2h_bytes},
            "gpu_launches": launches, "clocks": clocks,
        }
        line.update(wl.extra())
        if hasattr(wl, "b1_latency") and world == 1:
            line["b1_latency"] = wl.b1_latency()
        if world == 1 and not args.no_cpu_baseline:
            cw = type(wl).cpu_only(rank)
            cw.cpu_step(0)
            t0 = time.perf_counter()
            cw.cpu_step(1)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": cw.windows_per_step / dt, "unit": "windows/s", "cores": cw.cpu_cores,
                                    "kind": "port", "sample": cw.cpu_sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
