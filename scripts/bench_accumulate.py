"""Exploration benchmark of the L1 kernels (not the contract bench; see bench.py).
Times every accumulation entry point on BASELINE config 2 shapes with CUDA events, rotating
between input buffers larger than L2, and writes gpurun_out/accumulate_sweep.json."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evfly_b200 import _lib  # noqa: E402
from evfly_b200.events import L1, to_device  # noqa: E402
from evfly_b200.synthetic import synthetic_stream, synthetic_window  # noqa: E402

PEAK = 6552.0
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def time_it(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn(0)
    torch.cuda.synchronize()
    ts = []
    for i in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(i)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return {"us_med": ts[len(ts) // 2], "us_min": ts[0], "us_max": ts[-1]}


def main():
    H, W, B, n = 480, 640, 5, 10_000_000
    res = {"peak_gbs": PEAK, "cases": []}
    for dist in ("uniform", "clustered"):
        recs = [to_device(synthetic_window(s, n, H, W, distribution=dist)) for s in (0, 1)]  # 2 x 160 MB > L2
        counts = torch.zeros((2, H, W), dtype=torch.int32, device="cuda")
        voxel = torch.zeros((B, H, W), dtype=torch.float32, device="cuda")
        ws = L1.voxel_workspace(H, W, B, "cuda")

        def counts_only(i):
            counts.zero_()
            L1.accumulate_counts(recs[i & 1], H, W, out=counts, algo="red")

        def counts_binned(i):
            counts.zero_()
            L1.accumulate_counts(recs[i & 1], H, W, out=counts, algo="binned")

        def direct(i):
            counts.zero_(); voxel.zero_()
            L1.voxelize_window(recs[i & 1], H, W, B, 0, 33_333_333, counts=counts, voxel=voxel, algo=0)

        def staged(i):
            L1.voxelize_window(recs[i & 1], H, W, B, 0, 33_333_333, counts=counts, voxel=voxel, ws=ws, algo=1)

        def staged_voxel_only(i):
            L1.voxelize_window(recs[i & 1], H, W, B, 0, 33_333_333, counts=None, voxel=voxel, ws=ws, algo=1, want_counts=False)

        for name, fn, bytes_ in (("counts_only", counts_only, 16 * n + 2 * H * W * 4),
                                 ("counts_binned", counts_binned, 16 * n + 2 * H * W * 4),
                                 ("voxel_direct", direct, 16 * n + 7 * H * W * 4),
                                 ("voxel_staged", staged, 16 * n + 7 * H * W * 4),
                                 ("voxel_staged_noCounts", staged_voxel_only, 16 * n + 5 * H * W * 4)):
            t = time_it(fn)
            t.update(case=name, dist=dist, n_events=n, alg_bytes=bytes_,
                     gbs=bytes_ / t["us_med"] / 1e3, frac=bytes_ / t["us_med"] / 1e3 / PEAK,
                     gev_s=n / t["us_med"] / 1e3)
            print(json.dumps(t), flush=True)
            res["cases"].append(t)
        del recs

    # config 2b: 100 windows x 100k events
    T, per = 100, 100_000
    rec, edges = synthetic_stream(0, T, per, H, W)
    d = to_device(rec)
    d_edges = torch.from_numpy(edges).cuda()
    counts = torch.empty((T, 2, H, W), dtype=torch.int32, device="cuda")
    voxel = torch.empty((T, B, H, W), dtype=torch.float32, device="cuda")
    for name, vox in (("windows_counts+voxel", voxel), ("windows_counts", None)):
        def fn(i):
            L1.accumulate_windows(d, d_edges, H, W, B if vox is not None else None, counts=counts, voxel=vox)
        bytes_ = 16 * T * per + T * H * W * 4 * (2 + (B if vox is not None else 0))
        t = time_it(fn, iters=10)
        t.update(case=name, dist="uniform", n_events=T * per, alg_bytes=bytes_, gbs=bytes_ / t["us_med"] / 1e3,
                 frac=bytes_ / t["us_med"] / 1e3 / PEAK)
        print(json.dumps(t), flush=True)
        res["cases"].append(t)

    # batch-1 streaming window (config 5): 100k events at 480x640
    recs = [to_device(synthetic_window(s, 100_000, H, W)) for s in range(4)]
    counts1 = torch.zeros((2, H, W), dtype=torch.int32, device="cuda")
    voxel1 = torch.zeros((B, H, W), dtype=torch.float32, device="cuda")
    ws = L1.voxel_workspace(H, W, B, "cuda")
    t = time_it(lambda i: L1.voxelize_window(recs[i & 3], H, W, B, 0, 33_333_333, counts=counts1, voxel=voxel1, ws=ws, algo=1), iters=50)
    t.update(case="b1_window_100k_staged")
    print(json.dumps(t), flush=True)
    res["cases"].append(t)
    def b1_direct(i):
        counts1.zero_(); voxel1.zero_()
        L1.voxelize_window(recs[i & 3], H, W, B, 0, 33_333_333, counts=counts1, voxel=voxel1, algo=0)
    t = time_it(b1_direct, iters=50)
    t.update(case="b1_window_100k_direct")
    print(json.dumps(t), flush=True)
    res["cases"].append(t)

    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/accumulate_sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()
