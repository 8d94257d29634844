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

    # CPU port of the reference algorithm (oracle): one full window, single thread
    def cpu_step(self, i: int):
        from oracle import ev_oracle as O
        O.voxel_window(self.host[i & 1], self.H, self.W, self.B, self.T0, self.T1)
    cpu_sample = "1 window of 10M events (the full step), C loop restating node.cpp / histogram2d, 1 thread"
    cpu_cores = 1


WORKLOADS = {"accumulate": AccumulateWorkload}
try:
    from evfly_b200.bench_pipeline import PipelineWorkload  # added once the forward is on the path
    WORKLOADS["pipeline"] = PipelineWorkload
    DEFAULT_WORKLOAD = "pipeline"
except ImportError:
    DEFAULT_WORKLOAD = "accumulate"


# =============================================================================================
def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's algorithm on the host cores (rank 0 only)."""
    if rank != 0:
        return
    wl_cls = WORKLOADS[args.workload]
    wl = wl_cls.cpu_only(rank) if hasattr(wl_cls, "cpu_only") else None
    if wl is None:
        # accumulate workload needs no device objects for its CPU leg
        class _Shim(wl_cls):
            def __init__(self):
                from evfly_b200.synthetic import synthetic_window
                self.host = [synthetic_window(s, self.N_EV, self.H, self.W) for s in (0, 1)]
        wl = _Shim()
    for i in range(args.warmup):
        wl.cpu_step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        wl.cpu_step(i)
    dt = time.perf_counter() - t0
    val = wl.windows_per_step * args.steps / dt
    line = {"impl": "reference", "metric": "event_windows_per_sec", "value": val, "unit": "windows/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype,
            "data": "synthetic", "config": {"workload": wl.name},
            "cpu_baseline": {"value": val, "unit": "windows/s", "cores": wl.cpu_cores, "kind": "port", "sample": wl.cpu_sample},
            "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
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
        for i in range(warmup):
            fn(i)
        barrier()
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
    e2e_s, _, _ = timed(wl.e2e_step, max(3, args.steps // 2), 3)
    e2e_steps = max(3, args.steps // 2)

    if rank == 0:
        windows = wl.windows_per_step * world
        value = windows * args.steps / total_s
        line = {
            "metric": "event_windows_per_sec", "value": value, "unit": "windows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype,
            "data": "synthetic",
            "config": {"workload": wl.name, "windows_per_step_per_gpu": wl.windows_per_step,
                       "l2_policy": "inputs larger than L2 (rotating 2 x 160 MB event buffers)",
                       "sharding": "independent windows per rank, no data-path collective"},
            "roofline": wl.roofline(dom_s / args.steps, peaks),
            "e2e": {"value": windows * e2e_steps / e2e_s, "unit": "windows/s",
                    "h2d_bytes_per_step": wl.h2d_bytes, "d2h_bytes_per_step": wl.d2h_bytes},
            "gpu_launches": launches, "clocks": clocks,
        }
        if hasattr(wl, "extra"):
            line.update(wl.extra())
        if world == 1 and not args.no_cpu_baseline:
            t0 = time.perf_counter()
            wl.cpu_step(0)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": wl.windows_per_step / dt, "unit": "windows/s", "cores": wl.cpu_cores,
                                    "kind": "port", "sample": wl.cpu_sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
