"""Host-side enqueue time of one bench step (no synchronisation inside the loop) vs its GPU time."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = bench.WORKLOADS["trajectories"](0, torch.device("cuda", 0))
for i in range(3):
    wl.step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(5):
    wl.step(i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print({"host_enqueue_ms_per_step": (t1 - t0) / 5 * 1e3, "total_ms_per_step": (t2 - t0) / 5 * 1e3})
