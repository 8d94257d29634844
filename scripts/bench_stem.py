"""Time the UNet stem alone at bench size (400 frames 260x346x2)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evfly_b200 import tc
x = torch.rand(400, 2, 260, 346, device="cuda")
w = torch.randn(32, 2, 3, 3, device="cuda") * 0.3
b = torch.randn(32, device="cuda")
for fma in (False, True):
    for _ in range(3):
        tc.stem_conv3x3(x, w, b, fma=fma)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tc.stem_conv3x3(x, w, b, fma=fma)
    e1.record(); torch.cuda.synchronize()
    print("fma" if fma else "tc", os.environ.get("EVFLY_STEM_CTAS_PER_SM"), e0.elapsed_time(e1) / 10, "ms")
