"""One launch of the stem+e12 kernel (and of the two-kernel path it replaces) on a small batch: the command ncu wraps
for a `--set full` capture (scripts/gpu_batch_*.sh)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evfly_b200 import tc
torch.manual_seed(0)
N, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 260, 346
mask = (torch.rand(N, 1, H, W, device="cuda") < 0.35).float()
w1, b1 = torch.randn(32, 1, 3, 3, device="cuda") * 0.5, torch.randn(32, device="cuda") * 0.2
w2 = tc.pack_conv3x3_weight(torch.randn(32, 32, 3, 3, device="cuda") * 0.06)
b2 = torch.randn(32, device="cuda") * 0.1
for _ in range(3):
    out, pooled = tc.stem_e12_pool(mask, w1, b1, w2, b2)
    ref = tc.conv3x3_pool(tc.stem_conv3x3(mask, w1, b1), w2, b2)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(10):
    tc.stem_e12_pool(mask, w1, b1, w2, b2)
e[1].record()
for _ in range(10):
    tc.conv3x3_pool(tc.stem_conv3x3(mask, w1, b1), w2, b2)
e[2].record()
torch.cuda.synchronize()
print(f"N={N}: fused stem+e12+pool {e[0].elapsed_time(e[1]) / 10:.3f} ms, stem kernel then e12+pool {e[1].elapsed_time(e[2]) / 10:.3f} ms")
