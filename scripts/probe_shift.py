import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evfly_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
for KC in (32, 64):
    x = torch.randn(136, KC).to(torch.bfloat16).cuda()
    w = torch.randn(32, KC).to(torch.bfloat16).cuda()
    for shift in (0, 1, 2, 3, 7, 8):
        for ubo in (0, 1):
            out = torch.full((128, 32), float("nan"), device="cuda")
            _lib.check(lib.evfly_tc_shift_probe(x.data_ptr(), w.data_ptr(), out.data_ptr(), KC, shift, ubo, _lib.stream_ptr()))
            torch.cuda.synchronize()
            want = x[shift:shift + 128].float() @ w.float().t()
            err = (out - want).abs().max().item()
            # which row shift does the result correspond to, if any?
            best = min(range(0, 9), key=lambda s: (out - x[s:s + 128].float() @ w.float().t()).abs().max().item())
            print(f"KC={KC} shift={shift} base_offset={ubo}: max err {err:.4g}  (best-matching shift {best})")
