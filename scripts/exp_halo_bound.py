"""Elimination experiment (temporary debug switches in tc_conv_halo.cu): which role bounds k_tc_stem_e12 and the halo convs.
NOTE: the EVFLY_STEM_DBG / EVFLY_HALO_DBG switches this script drives were temporary instrumentation in tc_conv_halo.cu (roles reduced to
their barrier traffic, wait flavours); they are not in the shipped kernels. Results: profiles/r2_exp_halo_roles.txt."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evfly_b200 import tc
torch.manual_seed(0)
N, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 400, 260, 346
mask = (torch.rand(N, 1, H, W, device="cuda") < 0.1).float()
w1, b1 = torch.randn(32, 1, 3, 3, device="cuda") * 0.5, torch.randn(32, device="cuda") * 0.2
w2 = tc.pack_conv3x3_weight(torch.randn(32, 32, 3, 3, device="cuda") * 0.06)
b2 = torch.randn(32, device="cuda") * 0.1

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

os.environ["EVFLY_STEM_DBG"] = "0"
ref_out, ref_pool = tc.stem_e12_pool(mask, w1, b1, w2, b2)
names = {0: "try_wait (default)", 256: "test_wait spin", 512: "try_wait hint 32 ns", 768: "try_wait hint 256 ns", 1024: "try_wait hint 2000 ns"}
for d, nm in names.items():
    os.environ["EVFLY_STEM_DBG"] = str(d)
    ms = timeit(lambda: tc.stem_e12_pool(mask, w1, b1, w2, b2))
    o, pl = tc.stem_e12_pool(mask, w1, b1, w2, b2)
    print(f"  dbg={d:4d} {ms:8.3f} ms  {nm}  equal: {torch.equal(o.data, ref_out.data)} {torch.equal(pl.data, ref_pool.data)}", flush=True)
os.environ["EVFLY_STEM_DBG"] = "0"
os.environ.pop("EVFLY_STEM_2CTA", None)
sys.exit(0)
out, pooled = tc.stem_e12_pool(mask, w1, b1, w2, b2)
def layer(name, g, cin, cout, pool):
    w = tc.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device="cuda") * 0.05)
    b = torch.randn(cout, device="cuda") * 0.1
    print(f"{name}: N={g.N} valid {g.vh}x{g.vw} {cin}->{cout} pool={pool}")
    res = None
    for d, nm in {0: "baseline", 4: "1 MMA of 9*Cin/16", 2: "no global stores", 16: "no pool shuffles", 32: "epilogue = tmem_ld + arrive", 36: "1 MMA + trivial epilogue"}.items():
        os.environ["EVFLY_HALO_DBG"] = str(d)
        ms = timeit(lambda: tc.conv3x3_pool(g, w, b) if pool else tc.conv3x3(g, w, b))
        print(f"  dbg={d:4d} {ms:8.3f} ms  {nm}", flush=True)
    os.environ["EVFLY_HALO_DBG"] = "0"
    return tc.conv3x3_pool(g, w, b) if pool else tc.conv3x3(g, w, b)
e21 = layer("e21", pooled, 32, 64, False)
e22, p2 = layer("e22", e21, 64, 64, True)
e31 = layer("e31", p2, 64, 128, False)
e32, p3 = layer("e32", e31, 128, 128, True)
