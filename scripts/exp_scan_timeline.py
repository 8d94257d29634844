"""ConvLSTM recurrence at the bench shape (T = 100 steps, 16 trajectories x 204 pixels, 512 channels): the fused persistent kernel
against the x-gate GEMM + scan pair it replaces."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evfly_b200 import tc
torch.manual_seed(0)
T, n_traj, Ch = 100, 16, 512
P = n_traj * 204
dev = "cuda"
x = (torch.randn(T * P, Ch, device=dev)).to(torch.bfloat16)
wx = (torch.randn(4 * Ch, Ch, device=dev) * 0.02).to(torch.bfloat16)
wh = (torch.randn(4 * Ch, Ch, device=dev) * 0.02).to(torch.bfloat16)
h_all = torch.zeros((T + 1, P, Ch), dtype=torch.bfloat16, device=dev)
gx = torch.empty(T * P, 4 * Ch, device=dev)
c = torch.zeros(P, Ch, device=dev)

def timeit(fn, reps=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

t_gemm = timeit(lambda: tc.gemm(x, wx, None, out_f32=gx))
t_scan = timeit(lambda: tc.convlstm_scan(h_all, wh, gx, c, T, P, Ch))
t_fused = timeit(lambda: tc.convlstm_scan_fused(x, wx, h_all, wh, c, T, P, Ch))
print(f"T={T} P={P}: x-gate GEMM {t_gemm:.3f} ms + scan {t_scan:.3f} ms ({1e3 * t_scan / T:.1f} us/step) = {t_gemm + t_scan:.3f} ms; fused {t_fused:.3f} ms ({1e3 * t_fused / T:.1f} us/step)")
