"""Per-layer CUDA-event timing of the bf16 tensor-core conv kernel on the UNet's layer shapes."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evfly_b200 import tc

LAYERS = [  # name, Hp, Wp, vh_in, vw_in, Cin, Cout
    ("e12", 260, 346, 258, 344, 32, 32), ("e21", 128, 171, 128, 171, 32, 64), ("e22", 128, 171, 126, 169, 64, 64),
    ("e31", 62, 83, 62, 83, 64, 128), ("e32", 62, 83, 60, 81, 128, 128), ("e41", 29, 39, 29, 39, 128, 256),
    ("e42", 29, 39, 27, 37, 256, 256), ("e51", 12, 17, 12, 17, 256, 512), ("e52", 12, 17, 10, 15, 512, 512),
    ("d11", 16, 26, 16, 26, 512, 256), ("d12", 16, 26, 14, 24, 256, 256), ("d21", 24, 44, 24, 44, 256, 128),
    ("d22", 24, 44, 22, 42, 128, 128), ("d31", 40, 80, 40, 80, 128, 64), ("d32", 40, 80, 38, 78, 64, 64),
    ("d41", 72, 152, 72, 152, 64, 32), ("d42", 72, 152, 70, 150, 32, 32)]


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    iters = 5
    tot_t = tot_f = 0
    for name, Hp, Wp, vh, vw, Cin, Cout in LAYERS:
        g = tc.Grid(torch.randn((N, Hp, Wp, Cin), device="cuda").to(tc.BF16), vh, vw)
        w = tc.pack_conv3x3_weight(torch.randn((Cout, Cin, 3, 3), device="cuda") * (9 * Cin) ** -0.5)
        b = torch.randn(Cout, device="cuda")
        out = tc.new_grid(N, Hp, Wp, Cout, vh - 2, vw - 2, "cuda")
        for _ in range(2):
            tc.conv3x3(g, w, b, out=out)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            tc.conv3x3(g, w, b, out=out)
        e.record(); e.synchronize()
        us = a.elapsed_time(e) / iters * 1e3
        flops_valid = 2.0 * N * (vh - 2) * (vw - 2) * Cout * Cin * 9
        flops_grid = 2.0 * N * Hp * Wp * Cout * Cin * 9
        tot_t += us; tot_f += flops_valid
        print(json.dumps(dict(layer=name, N=N, us=round(us, 1), tflops_valid=round(flops_valid / us / 1e6, 1), tflops_grid=round(flops_grid / us / 1e6, 1),
                              tiles=N * Hp * Wp // 128, us_per_tile_per_sm=round(us / max(1, (N * Hp * Wp // 128) / 148), 3))), flush=True)
    print(json.dumps(dict(total_us=round(tot_t, 1), tflops_valid=round(tot_f / tot_t / 1e6, 1))))


if __name__ == "__main__":
    main()
