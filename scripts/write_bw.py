"""Pure-write and pure-read HBM bandwidth next to the driver's copy figure (MEASURED_PEAKS.json hbm_gbs = read + write of a
copy): what a write-dominated kernel (cfg 2b: 860 MB of frames out for 160 MB of events in) can reach at best."""
import json, torch
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3
out = {"bytes": n,
       "memset_write_gbs": n / t(lambda: a.zero_()) / 1e9,
       "copy_read_plus_write_gbs": 2 * n / t(lambda: b.copy_(a)) / 1e9,
       "reduce_read_gbs": n / t(lambda: a.view(torch.int32).sum()) / 1e9}
print(json.dumps(out))
