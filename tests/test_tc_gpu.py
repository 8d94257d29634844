"""GPU: the tcgen05/TMEM/TMA implicit-GEMM kernel and the NHWC glue kernels of the bf16 path,
against PyTorch on the CPU evaluated on the SAME bf16-rounded operands (fp64 accumulate), so the
only differences are fp32 accumulation order and the final bf16 rounding of the output."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from evfly_b200 import tc

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def bf(x):
    return x.to(BF).float()


def check_bf16(got, want, what=""):
    got, want = got.detach().float().cpu().double(), want.double()
    err = (got - want).abs()
    tol = 2 ** -7 * want.abs() + 2e-2          # bf16 output rounding (2^-8 rel) + accumulation slack
    assert (err <= tol).all(), f"{what}: max err {err.max():.4g} at |want| {want.abs().flatten()[err.argmax()]:.4g}"
    # and on average it is much tighter than the bound
    assert err.mean() <= 2 ** -8 * want.abs().mean() + 1e-3, f"{what}: mean err {err.mean():.4g}"


@pytest.mark.parametrize("M,K,N", [(300, 64, 64), (128, 32, 32), (1000, 512, 2048), (77, 96, 40), (4096, 256, 256), (204, 512, 2048)])
def test_tc_gemm(cuda_lib, M, K, N):
    x, w, b = bf(rnd(M, K, seed=1)), bf(rnd(N, K, seed=2, scale=K ** -0.5)), rnd(N, seed=3)
    want = x.double() @ w.double().t() + b.double()
    got = tc.gemm(x.to(BF).cuda(), w.to(BF).cuda(), b.cuda())
    check_bf16(got, want, "gemm")
    out32 = torch.empty((M, N), device="cuda")
    res = rnd(M, N, seed=4)
    tc.gemm(x.to(BF).cuda(), w.to(BF).cuda(), None, relu=True, out_f32=out32, res_f32=res.cuda())
    want = F.relu(x.double() @ w.double().t() + res.double())
    np.testing.assert_allclose(out32.cpu().double().numpy(), want.numpy(), rtol=1e-4, atol=1e-3)   # fp32 out: accumulation only


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 20, 24, 32, 32), (1, 31, 17, 64, 64), (2, 12, 17, 64, 128), (1, 29, 39, 128, 128),
                                            (1, 12, 17, 256, 512), (1, 12, 17, 512, 512), (3, 14, 24, 512, 256), (1, 40, 50, 128, 64)])
def test_tc_conv3x3_valid(cuda_lib, N, H, W, Cin, Cout):
    x = bf(rnd(N, Cin, H, W, seed=1))
    w = bf(rnd(Cout, Cin, 3, 3, seed=2, scale=(9 * Cin) ** -0.5))
    b = rnd(Cout, seed=3)
    want = F.relu(F.conv2d(x.double(), w.double(), b.double()))                   # [N,Cout,H-2,W-2]
    g = tc.nchw_to_grid(x.cuda(), H, W)
    out = tc.conv3x3(g, tc.pack_conv3x3_weight(w.cuda()), b.cuda(), relu=True)
    assert (out.vh, out.vw) == (H - 2, W - 2) and out.data.shape == (N, H, W, Cout)
    check_bf16(tc.grid_to_nchw(out.data, H - 2, W - 2), want, "conv3x3")
    # second conv on the same pitch grid: the don't-care border of the first must not leak
    w2 = bf(rnd(Cout, Cout, 3, 3, seed=5, scale=(9 * Cout) ** -0.5))
    want2 = F.conv2d(bf(want.float()).double(), w2.double())
    out2 = tc.conv3x3(out, tc.pack_conv3x3_weight(w2.cuda()), None, relu=False)
    check_bf16(tc.grid_to_nchw(out2.data, H - 4, W - 4), want2, "conv3x3 x2")


@pytest.mark.parametrize("Cin,Cout,vh,vw,Hp,Wp", [(512, 256, 8, 13, 12, 17), (64, 32, 84, 128, 88, 132), (128, 64, 5, 7, 5, 7)])
def test_tc_conv_transpose(cuda_lib, Cin, Cout, vh, vw, Hp, Wp):
    N = 2
    m = torch.nn.ConvTranspose2d(Cin, Cout, 2, 2)
    x = bf(rnd(N, Cin, vh, vw, seed=1))
    w = bf(m.weight.detach())
    want = F.conv_transpose2d(x.double(), w.double(), m.bias.detach().double(), stride=2)
    g = tc.nchw_to_grid(x.cuda(), Hp, Wp)
    cat = torch.zeros((N, 2 * vh, 2 * vw, 2 * Cout), dtype=BF, device="cuda")
    tc.conv_transpose2x2(g, tc.pack_convt2x2_weight(w.cuda()), m.bias.detach().cuda(), cat, Cout)
    check_bf16(tc.grid_to_nchw(cat, 2 * vh, 2 * vw)[:, Cout:], want, "convT")
    assert not cat[..., :Cout].any()


def test_nhwc_glue(cuda_lib):
    N, C, H, W = 2, 64, 25, 35
    x = bf(rnd(N, C, H, W, seed=1))
    g = tc.nchw_to_grid(x.cuda(), H + 3, W + 2)
    assert torch.equal(tc.grid_to_nchw(g.data, H, W).cpu(), x)
    p = tc.maxpool2x2(g)
    assert torch.equal(tc.grid_to_nchw(p.data, H // 2, W // 2).cpu(), F.max_pool2d(x, 2))
    cat = torch.zeros((N, 16, 26, 2 * C), dtype=BF, device="cuda")
    tc.resize_bilinear_into(g, 16, 26, cat, 0)
    want = F.interpolate(x, size=(16, 26), mode="bilinear", align_corners=False)
    check_bf16(tc.grid_to_nchw(cat, 16, 26)[:, :C], want, "bilinear")
    tc.crop_into(g, 4, 4, 16, 26, cat, C)
    assert torch.equal(tc.grid_to_nchw(cat, 16, 26)[:, C:].cpu(), x[:, :, 4:20, 4:30])
    # stem: fp32 NCHW (1 or 2 channels) -> bf16 NHWC 32
    for cin in (1, 2):
        xs, ws, bs = rnd(2, cin, 30, 41, seed=2), rnd(32, cin, 3, 3, seed=3, scale=0.3), rnd(32, seed=4)
        out = tc.stem_conv3x3(xs.cuda(), ws.cuda(), bs.cuda())
        check_bf16(tc.grid_to_nchw(out.data, 28, 39), F.relu(F.conv2d(xs, ws, bs)), "stem")
    # ConvLSTM cell update
    P, Ch = 204, 512
    gates, c = rnd(P, 4 * Ch, seed=5), rnd(P, Ch, seed=6)
    i, f, o, gg = gates.split(Ch, dim=1)
    cn = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    hn = torch.sigmoid(o) * torch.tanh(cn)
    dc, dh = c.clone().cuda(), torch.empty((P, Ch), dtype=BF, device="cuda")
    tc.convlstm_pointwise(gates.cuda(), dc, dh)
    np.testing.assert_allclose(dc.cpu().numpy(), cn.numpy(), rtol=1e-5, atol=1e-5)
    check_bf16(dh, hn, "convlstm h")
