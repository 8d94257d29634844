"""Run the binned count accumulation a few times (for ncu captures)."""
import torch
from evfly_b200.events import L1, to_device
from evfly_b200.synthetic import synthetic_window
H, W, n = 480, 640, 10_000_000
recs = [to_device(synthetic_window(s, n, H, W)) for s in (0, 1)]
counts = torch.zeros((2, H, W), dtype=torch.int32, device="cuda")
for i in range(4):
    counts.zero_()
    L1.accumulate_counts(recs[i & 1], H, W, out=counts, algo="binned")
torch.cuda.synchronize()
print(int(counts.sum()))
