"""Timeline of CTA 0 of k_tc_stem_e12 (temporary instrumentation): clock64 at the role hand-offs of its first 64 tiles.
NOTE: needs the temporary clock64 instrumentation of k_tc_stem_e12 (EVFLY_STEM_TL), which is not in the shipped kernel.
Results: profiles/r2_exp_halo_roles.txt (timeline paragraph)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evfly_b200 import tc
torch.manual_seed(0)
N, H, W = 400, 260, 346
mask = (torch.rand(N, 1, H, W, device="cuda") < 0.1).float()
w1, b1 = torch.randn(32, 1, 3, 3, device="cuda") * 0.5, torch.randn(32, device="cuda") * 0.2
w2 = tc.pack_conv3x3_weight(torch.randn(32, 32, 3, 3, device="cuda") * 0.06)
b2 = torch.randn(32, device="cuda") * 0.1
for _ in range(3): tc.stem_e12_pool(mask, w1, b1, w2, b2)
tl = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
os.environ["EVFLY_STEM_TL"] = str(tl.data_ptr())
tc.stem_e12_pool(mask, w1, b1, w2, b2)
torch.cuda.synchronize()
t = tl.view(64, 16).cpu()
t0 = int(t[t > 0].min())
print("local tile: producer got stage | producer arrived full | MMA past tempty | MMA past full | MMAs issued+committed | epilogue woke | epilogue arrived tempty   (clk since first event)")
for i in range(40):
    r = [int(x) - t0 if x > 0 else -1 for x in t[i]]
    print(f"{i:3d}: P {r[0]:7d} {r[1]:7d} | M {r[7]:7d} {r[2]:7d} {r[3]:7d} | E woke {r[4]:7d} tmem {r[8] - r[4]:5d} math+sts {r[9] - r[8]:5d} pool {r[10] - r[9]:5d} writeout {r[11] - r[10]:5d} arrive {r[5] - r[11]:5d}")
