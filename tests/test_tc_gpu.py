"""GPU: the tcgen05/TMEM/TMA implicit-GEMM kernel and the NHWC glue kernels of the bf16 path,
against PyTorch on the CPU evaluated on the SAME bf16-rounded operands (fp64 accumulate), so the
only differences are fp32 accumulation order and the final bf16 rounding of the output."""
import os

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
        # the tensor-core stem rounds image and weights to bf16 (bias stays fp32)
        check_bf16(tc.grid_to_nchw(out.data, 28, 39), F.relu(F.conv2d(bf(xs).double(), bf(ws).double(), bs.double())), "stem")
        fma = tc.stem_conv3x3(xs.cuda(), ws.cuda(), bs.cuda(), fma=True)     # fp32 CUDA-core version of the same layer
        check_bf16(tc.grid_to_nchw(fma.data, 28, 39), F.relu(F.conv2d(xs, ws, bs)), "stem fma")
        d = (tc.grid_to_nchw(out.data, 28, 39) - tc.grid_to_nchw(fma.data, 28, 39)).abs().max().item()
        assert d <= 0.05, d
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


# ---- bf16 ViT kernels ---------------------------------------------------------------------------
@pytest.mark.parametrize("B", [2, 30])       # 30: >= 1024 tokens, the tensor-core kernels (operands rounded to bf16)
def test_patch_embed_ln_and_reduction_conv(cuda_lib, B):
    rt = bf if B >= 30 else (lambda t: t)
    # stage 1: fp32 depth image, 7x7 stride 4 pad 3, 1 -> 32, LayerNorm
    x = rnd(B, 1, 60, 90, seed=1)
    w, b, g, be = rnd(32, 1, 7, 7, seed=2, scale=0.15), rnd(32, seed=3), 1 + 0.1 * rnd(32, seed=4), 0.1 * rnd(32, seed=5)
    y = F.conv2d(rt(x), rt(w), b, stride=4, padding=3)
    want = F.layer_norm(y.flatten(2).transpose(1, 2), (32,), g, be)
    tok, H, W = tc.patch_embed_ln(x.cuda(), True, tc.pack_conv_kc(w.cuda()), b.cuda(), g.cuda(), be.cuda(), B, 60, 90, 1, 32, 7, 4, 3, 1e-5)
    assert (H, W) == (15, 23)
    check_bf16(tok, want, "patch embed 1")
    # stage 2 and the reduction convs: bf16 NHWC input
    for Cin, Cout, k, s, p, Hh, Ww in [(32, 64, 3, 2, 1, 15, 23), (32, 32, 8, 8, 0, 15, 23), (64, 64, 4, 4, 0, 8, 12)]:
        xt = bf(rnd(B, Hh * Ww, Cin, seed=6))
        w, b = rnd(Cout, Cin, k, k, seed=7, scale=(Cin * k * k) ** -0.5), rnd(Cout, seed=8)
        g, be = 1 + 0.1 * rnd(Cout, seed=9), 0.1 * rnd(Cout, seed=10)
        y = F.conv2d(xt.view(B, Hh, Ww, Cin).permute(0, 3, 1, 2), rt(w) if k == 3 else w, b, stride=s, padding=p)
        want = F.layer_norm(y.flatten(2).transpose(1, 2), (Cout,), g, be)
        tok, H, W = tc.patch_embed_ln(xt.to(BF).cuda(), False, tc.pack_conv_kc(w.cuda()), b.cuda(), g.cuda(), be.cuda(), B, Hh, Ww, Cin, Cout, k, s, p, 1e-5)
        assert tok.shape == want.shape
        check_bf16(tok, want, f"patch embed {Cin}->{Cout} k{k}")


def test_layernorm_attention_dwconv_bf16(cuda_lib):
    for C in (32, 64):
        x, g, b = bf(rnd(500, C, seed=1) * 2 + 0.5), rnd(C, seed=2), rnd(C, seed=3)
        check_bf16(tc.layernorm_bf16(x.to(BF).cuda(), g.cuda(), b.cuda(), 1e-5), F.layer_norm(x, (C,), g, b), "layernorm")
    for B, N, C, heads, nkv in [(3, 345, 32, 1, 2), (3, 96, 64, 2, 6)]:
        q, kv = bf(rnd(B, N, C, seed=1)), bf(rnd(B, nkv, 2 * C, seed=2))
        d = C // heads
        kvr = kv.reshape(B, nkv, 2, heads, d).permute(2, 0, 3, 1, 4)
        qr = q.reshape(B, N, heads, d).permute(0, 2, 1, 3)
        att = torch.softmax(qr @ kvr[0].transpose(-2, -1) / d ** 0.5, dim=-1)
        want = (att @ kvr[1]).transpose(1, 2).reshape(B, N, C)
        check_bf16(tc.attention_small_bf16(q.to(BF).cuda(), kv.to(BF).cuda(), heads), want, "attention")
    # the last three shapes take the 4-pixels-per-thread kernel (>= 4096 pixels), incl. a row width that is not a multiple of 4
    for B, H, W, C in [(2, 15, 23, 32), (2, 8, 12, 64), (16, 15, 23, 32), (48, 8, 12, 64), (5, 30, 37, 16)]:
        Ce = 8 * C
        x = bf(rnd(B, H, W, Ce, seed=1))
        w, b = rnd(Ce, 8, 3, 3, seed=2, scale=72 ** -0.5), rnd(Ce, seed=3)
        want = F.gelu(F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=1, groups=C)).permute(0, 2, 3, 1)
        check_bf16(tc.dwconv3x3_gelu(x.to(BF).cuda(), w.cuda(), b.cuda()), want, "dwconv+gelu")


def test_gemm_tokens_residual_and_small_m_tiles(cuda_lib):
    M, K, N = 1035, 256, 32
    x, w, b, r = bf(rnd(M, K, seed=1)), bf(rnd(N, K, seed=2, scale=K ** -0.5)), rnd(N, seed=3), bf(rnd(M, N, seed=4))
    got = tc.gemm_tokens(x.to(BF).cuda(), w.to(BF).cuda(), b.cuda(), res_bf16=r.to(BF).cuda())
    check_bf16(got, x.double() @ w.double().t() + b.double() + r.double(), "gemm + bf16 residual")
    # small M, wide N: the launcher narrows the N tile to fill the SMs (ConvLSTM step shape)
    M, K, N = 204, 512, 2048
    x, w = bf(rnd(M, K, seed=5)), bf(rnd(N, K, seed=6, scale=K ** -0.5))
    check_bf16(tc.gemm(x.to(BF).cuda(), w.to(BF).cuda()), x.double() @ w.double().t(), "small-M gemm")


def test_lstm_smem_weights(cuda_lib):
    from evfly_b200 import ops
    from evfly_b200._modbase import pack_lstm, run_lstm
    torch.manual_seed(0)
    m = torch.nn.LSTM(input_size=517, hidden_size=128, num_layers=3).eval()
    x = rnd(40, 517, seed=1)
    h0, c0 = rnd(3, 128, seed=2) * 0.3, rnd(3, 128, seed=3) * 0.3
    with torch.no_grad():
        want, (hn, cn) = m(x, (h0, c0))
        mc = m.cuda()
        got, (h, c) = run_lstm(ops, pack_lstm(mc), x.cuda(), (h0, c0), 128, smem_weights=True)
    # W_hh is rounded to bf16 (2^-9 relative), everything else fp32
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-2, atol=3e-3)
    np.testing.assert_allclose(h.cpu().numpy(), hn.numpy(), rtol=2e-2, atol=3e-3)
    np.testing.assert_allclose(c.cpu().numpy(), cn.numpy(), rtol=2e-2, atol=3e-3)


def test_fused_convlstm_step(cuda_lib):
    P, Ch = 204, 512
    h_prev, wh = bf(rnd(P, Ch, seed=1) * 0.5), bf(rnd(4 * Ch, Ch, seed=2, scale=Ch ** -0.5))
    gx, c = rnd(P, 4 * Ch, seed=3), rnd(P, Ch, seed=4)
    gates = h_prev.double() @ wh.double().t() + gx.double()
    i, f, o, g = gates.split(Ch, dim=1)                       # gate-major (convlstm.py:44)
    cn = torch.sigmoid(f) * c.double() + torch.sigmoid(i) * torch.tanh(g)
    hn = torch.sigmoid(o) * torch.tanh(cn)
    # device: rows / columns interleaved n = 4*ch + gate
    inter = lambda t2d: t2d.reshape(-1, 4, Ch).permute(0, 2, 1).reshape(-1, 4 * Ch).contiguous()
    dc, dh = c.clone().cuda(), torch.empty((P, Ch), dtype=BF, device="cuda")
    tc.convlstm_step(h_prev.to(BF).cuda(), tc.pack_convlstm_gate_weight(wh.cuda()), inter(gx).cuda(), dc, dh)
    np.testing.assert_allclose(dc.cpu().double().numpy(), cn.numpy(), rtol=2e-3, atol=2e-3)
    check_bf16(dh, hn, "fused convlstm h")


def test_shifted_descriptor_probe(cuda_lib):
    """The hardware assumption of the halo-reuse conv: a UMMA descriptor may start at any row of a TMA-written
    swizzled tile (base_offset = 0, swizzle on absolute address bits)."""
    import ctypes
    from evfly_b200 import _lib
    probe_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "libevfly_tc_probe.so")
    assert os.path.exists(probe_path), "tests/native/libevfly_tc_probe.so is built by __graft_entry__.build()"
    probe = ctypes.CDLL(probe_path).evfly_tc_shift_probe
    probe.restype, probe.argtypes = ctypes.c_int, [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
    for KC in (32, 64):
        x, w = rnd(136, KC, seed=1).to(BF).cuda(), rnd(32, KC, seed=2).to(BF).cuda()
        for shift in (0, 1, 2, 5, 8):
            out = torch.empty((128, 32), device="cuda")
            _lib.check(probe(x.data_ptr(), w.data_ptr(), out.data_ptr(), KC, shift, 0, _lib.stream_ptr()))
            want = x[shift:shift + 128].float() @ w.float().t()
            assert torch.allclose(out, want, rtol=1e-4, atol=1e-4), (KC, shift)


@pytest.mark.parametrize("N,H,W,vh,vw,Cin,Cout", [(2, 20, 24, 20, 24, 32, 32), (1, 37, 29, 35, 27, 64, 64), (3, 19, 11, 19, 11, 32, 64),
                                                  (1, 72, 152, 72, 152, 64, 32), (2, 18, 10, 18, 10, 32, 32), (1, 5, 5, 3, 3, 64, 64),
                                                  (2, 41, 53, 40, 50, 64, 128), (1, 126, 170, 126, 169, 64, 128),
                                                  # Cin = 128: halo reuse with the weights streamed through a ring
                                                  (2, 41, 53, 40, 50, 128, 128), (1, 62, 83, 62, 83, 128, 64), (3, 31, 41, 29, 39, 128, 256), (5, 20, 19, 20, 19, 128, 128)])
def test_halo_conv_equals_reference_and_streaming_kernel(cuda_lib, N, H, W, vh, vw, Cin, Cout):
    x = bf(rnd(N, Cin, vh, vw, seed=1))
    w = bf(rnd(Cout, Cin, 3, 3, seed=2, scale=(9 * Cin) ** -0.5))
    b = rnd(Cout, seed=3)
    want = F.relu(F.conv2d(x.double(), w.double(), b.double()))
    g = tc.nchw_to_grid(x.cuda(), H, W)
    wp = tc.pack_conv3x3_weight(w.cuda())
    out = tc.conv3x3(g, wp, b.cuda(), relu=True)
    got = tc.grid_to_nchw(out.data, vh - 2, vw - 2)
    check_bf16(got, want, "halo conv")
    tc.USE_HALO = False
    try:
        ref = tc.grid_to_nchw(tc.conv3x3(g, wp, b.cuda(), relu=True).data, vh - 2, vw - 2)
    finally:
        tc.USE_HALO = True
    assert (got - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()      # same math, different accumulation order


@pytest.mark.parametrize("N,cin,H,W", [(1, 2, 260, 346), (3, 1, 17, 23), (5, 2, 3, 3), (2, 2, 64, 129)])
def test_stem_tensor_core_matches_reference(cuda_lib, N, cin, H, W):
    """learner_models.py OrigUNet unet_e11 + ReLU. Sizes cover partial last tiles and rows that straddle tiles."""
    xs, ws, bs = rnd(N, cin, H, W, seed=12), rnd(32, cin, 3, 3, seed=13, scale=0.3), rnd(32, seed=14)
    out = tc.stem_conv3x3(xs.cuda(), ws.cuda(), bs.cuda())
    assert (out.vh, out.vw) == (H - 2, W - 2)
    check_bf16(tc.grid_to_nchw(out.data, H - 2, W - 2), F.relu(F.conv2d(bf(xs).double(), bf(ws).double(), bs.double())), "stem tc")


@pytest.mark.parametrize("N,vh,vw,Cin,Cout", [(2, 20, 24, 32, 32), (1, 37, 29, 64, 64), (3, 19, 13, 32, 64), (1, 71, 150, 64, 128), (1, 4, 4, 32, 32), (2, 60, 81, 128, 128)])
def test_halo_conv_fused_maxpool_is_bit_identical_to_separate_pool(cuda_lib, N, vh, vw, Cin, Cout):
    """learner_models.py OrigUNet: pool(relu(conv(x))). The fused epilogue pools fp32 values before rounding to bf16;
    rounding is monotonic, so the result must equal the stand-alone pool of the stored bf16 tensor bit for bit."""
    x = bf(rnd(N, Cin, vh, vw, seed=4))
    w = bf(rnd(Cout, Cin, 3, 3, seed=5, scale=(9 * Cin) ** -0.5))
    b = rnd(Cout, seed=6)
    g = tc.nchw_to_grid(x.cuda(), vh, vw)
    wp = tc.pack_conv3x3_weight(w.cuda())
    y, pooled = tc.conv3x3_pool(g, wp, b.cuda(), relu=True)
    assert (pooled.vh, pooled.vw) == ((vh - 2) // 2, (vw - 2) // 2)
    y2 = tc.conv3x3(g, wp, b.cuda(), relu=True)
    assert torch.equal(tc.grid_to_nchw(y.data, vh - 2, vw - 2), tc.grid_to_nchw(y2.data, vh - 2, vw - 2))
    want = tc.maxpool2x2(y2)
    assert torch.equal(tc.grid_to_nchw(pooled.data, pooled.vh, pooled.vw), tc.grid_to_nchw(want.data, want.vh, want.vw))
    ref = F.max_pool2d(F.relu(F.conv2d(x.double(), w.double(), b.double())), 2)
    check_bf16(tc.grid_to_nchw(pooled.data, pooled.vh, pooled.vw), ref, "fused pool")


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(3, 12, 17, 256, 512), (2, 29, 39, 256, 256), (1, 16, 26, 512, 256), (5, 5, 7, 256, 256),
                                            (2, 40, 51, 64, 64), (1, 33, 20, 32, 64), (2, 29, 39, 128, 256), (1, 24, 44, 128, 128), (2, 19, 27, 64, 32)])
def test_generic_conv_compact_output(cuda_lib, N, H, W, Cin, Cout):
    """EVFLY_TC_COMPACT / evfly_tc_conv3x3_halo_compact_bf16: the convs write their result on a grid whose pitch is the valid extent; the valid pixels are those of
    the pitch-preserving call bit for bit, and a second conv on the compact grid equals the second conv on the pitch grid."""
    x = bf(rnd(N, Cin, H, W, seed=1))
    w = bf(rnd(Cout, Cin, 3, 3, seed=2, scale=(9 * Cin) ** -0.5))
    b = rnd(Cout, seed=3)
    g = tc.nchw_to_grid(x.cuda(), H, W)
    wp = tc.pack_conv3x3_weight(w.cuda())
    y_pitch = tc.conv3x3(g, wp, b.cuda(), relu=True)
    y_comp = tc.conv3x3(g, wp, b.cuda(), relu=True, compact=True)
    assert (y_comp.Hp, y_comp.Wp, y_comp.vh, y_comp.vw) == (H - 2, W - 2, H - 2, W - 2) and (y_pitch.Hp, y_pitch.Wp) == (H, W)
    assert torch.equal(y_comp.data, y_pitch.data[:, :H - 2, :W - 2].contiguous())
    if H >= 7 and Cout % 32 == 0:
        w2 = tc.pack_conv3x3_weight(bf(rnd(Cout, Cout, 3, 3, seed=5, scale=(9 * Cout) ** -0.5)).cuda())
        z_pitch = tc.conv3x3(y_pitch, w2, None, relu=False)
        z_comp = tc.conv3x3(y_comp, w2, None, relu=False, compact=True)
        assert torch.equal(z_comp.data, z_pitch.data[:, :H - 4, :W - 4].contiguous())


@pytest.mark.parametrize("N,vh,vw,Cin,Cout,OH,OW", [(2, 130, 173, 64, 64, 34, 74), (1, 62, 83, 128, 128, 14, 34), (2, 40, 50, 32, 32, 7, 9),
                                                    (1, 37, 29, 64, 64, 35, 27), (1, 40, 33, 32, 64, 76, 60), (3, 21, 17, 64, 64, 1, 1)])
def test_pool_conv_writing_only_the_rows_the_skip_reads(cuda_lib, N, vh, vw, Cin, Cout, OH, OW):
    """learner_models.py:512-519: with the 'interp' skip y_e{k} is read only by F.interpolate(..., bilinear) to the decoder's size.
    The _rows variant of the fused-pool conv must leave that resize (and the pooled tensor) bit-identical while skipping the
    rows it does not sample; down- and up-sampling ratios, OH = 1 and ratios that land on row boundaries are covered."""
    x = bf(rnd(N, Cin, vh, vw, seed=4))
    w = bf(rnd(Cout, Cin, 3, 3, seed=5, scale=(9 * Cin) ** -0.5))
    b = rnd(Cout, seed=6)
    g = tc.nchw_to_grid(x.cuda(), vh, vw)
    wp = tc.pack_conv3x3_weight(w.cuda())
    y, pooled = tc.conv3x3_pool(g, wp, b.cuda(), relu=True)
    cat_full = torch.zeros((N, OH, OW, 2 * Cout), dtype=BF, device="cuda")
    tc.resize_bilinear_into(y, OH, OW, cat_full, 0)
    # poison the output buffer the sparse variant writes into, so that a row it wrongly skips shows
    import evfly_b200.tc as T_
    orig_new = T_.new_grid
    def poisoned(*a, **k):
        gg = orig_new(*a, **k)
        gg.data.fill_(float("nan"))
        return gg
    T_.new_grid = poisoned
    try:
        y2, pooled2 = tc.conv3x3_pool(g, wp, b.cuda(), relu=True, skip_rows=OH)
    finally:
        T_.new_grid = orig_new
    cat_rows = torch.zeros_like(cat_full)
    tc.resize_bilinear_into(y2, OH, OW, cat_rows, 0)
    assert torch.equal(cat_rows, cat_full)
    assert torch.equal(tc.grid_to_nchw(pooled2.data, pooled2.vh, pooled2.vw), tc.grid_to_nchw(pooled.data, pooled.vh, pooled.vw))
    written = ~torch.isnan(y2.data[:, :y2.vh, :y2.vw].float()).all(dim=(0, 2, 3))
    if OH * 3 < y2.vh:
        assert written.sum().item() <= 2 * OH + 2 and written.sum().item() < y2.vh        # it really skipped rows


@pytest.mark.parametrize("T,P,Cx,Ch", [(12, 204, 512, 512), (6, 3264, 512, 512), (9, 816, 256, 512), (3, 100, 64, 64), (5, 5000, 512, 512)])
def test_fused_convlstm_scan(cuda_lib, T, P, Cx, Ch):
    """convlstm.py:44-53 with the x half of the gate conv inside the persistent step kernel (csrc/convlstm_scan.cu): against the
    fp64 recurrence on the bf16-rounded operands, and against the x-gate GEMM + scan pair it replaces (same math, the x-gates are
    not rounded to fp32 in between). P = 3264 is the bench shape (two tiles on some CTAs), P = 5000 needs several rounds per step."""
    wx = tc.pack_convlstm_gate_weight((rnd(4 * Ch, Cx, seed=1, scale=Cx ** -0.5)).cuda())
    wh = tc.pack_convlstm_gate_weight((rnd(4 * Ch, Ch, seed=2, scale=Ch ** -0.5)).cuda())
    x = (rnd(T * P, Cx, seed=3)).to(BF).cuda()
    h0 = (rnd(P, Ch, seed=4) * 0.3).to(BF).cuda()
    c0 = (rnd(P, Ch, seed=5) * 0.5).cuda()
    h_all = torch.empty((T + 1, P, Ch), dtype=BF, device="cuda")
    h_all[0].copy_(h0)
    c = c0.clone()
    assert tc.convlstm_scan_fused(x, wx, h_all, wh, c, T, P, Ch)
    # the pair it replaces
    h_ref = torch.empty_like(h_all)
    h_ref[0].copy_(h0)
    c_ref = c0.clone()
    gx = torch.empty((T * P, 4 * Ch), device="cuda")
    tc.gemm(x, wx, None, out_f32=gx)
    tc.convlstm_scan(h_ref, wh, gx, c_ref, T, P, Ch)
    torch.cuda.synchronize()
    dh = (h_all.float() - h_ref.float()).abs().max().item()
    assert dh <= 2 ** -6 and (c - c_ref).abs().max().item() <= 1e-3, (dh, (c - c_ref).abs().max().item())      # summation order + bf16 rounding flips of h
    if T * P * Ch <= 12 * 816 * 512:
        wxd, whd = wx.float().cpu().double(), wh.float().cpu().double()
        h, cc = h0.float().cpu().double(), c0.cpu().double()
        xd = x.float().cpu().double()
        for t in range(T):
            g = (xd[t * P:(t + 1) * P] @ wxd.t() + h @ whd.t()).view(P, Ch, 4)
            cc = torch.sigmoid(g[..., 1]) * cc + torch.sigmoid(g[..., 0]) * torch.tanh(g[..., 3])
            h = (torch.sigmoid(g[..., 2]) * torch.tanh(cc)).to(BF).double()
        check_bf16(h_all[T], h, "fused scan h_T")
        np.testing.assert_allclose(c.cpu().double().numpy(), cc.numpy(), rtol=0, atol=2e-2)


@pytest.mark.parametrize("T,P,Ch", [(12, 204, 512), (7, 816, 512), (3, 100, 64)])
def test_persistent_convlstm_scan_equals_per_step_launches(cuda_lib, T, P, Ch):
    wh = tc.pack_convlstm_gate_weight((rnd(4 * Ch, Ch, seed=1, scale=Ch ** -0.5)).cuda())
    gx = rnd(T * P, 4 * Ch, seed=2).cuda()
    h0 = (rnd(P, Ch, seed=3) * 0.3).to(BF).cuda()
    outs = []
    for persistent in (True, False):
        tc.PERSISTENT_SCAN = persistent
        try:
            h_all = torch.empty((T + 1, P, Ch), dtype=BF, device="cuda")
            h_all[0].copy_(h0)
            c = torch.zeros((P, Ch), device="cuda")
            tc.convlstm_scan(h_all, wh, gx, c, T, P, Ch)
            torch.cuda.synchronize()
            outs.append((h_all.clone(), c.clone()))
        finally:
            tc.PERSISTENT_SCAN = True
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])     # same arithmetic, bit-identical
    # and both follow the recurrence (fp64 reference on the bf16-rounded operands)
    w = wh.float().cpu().double()
    h, cc = h0.float().cpu().double(), torch.zeros(P, Ch, dtype=torch.float64)
    for t in range(T):
        g = (h @ w.t() + gx[t * P:(t + 1) * P].cpu().double()).view(P, Ch, 4)
        cc = torch.sigmoid(g[..., 1]) * cc + torch.sigmoid(g[..., 0]) * torch.tanh(g[..., 3])
        h = (torch.sigmoid(g[..., 2]) * torch.tanh(cc)).to(BF).double()
    check_bf16(outs[0][0][T], h, "scan h_T")


def test_vit_tail_cat_and_same_padding_conv(cuda_lib):
    """vitfly_models.py:136-142: cat([pxShuffle(s2), up_sample(s1)]) -> down_sample(k=3, padding=1), bf16 NHWC on the
    tensor cores (channels zero-padded 48->64, 12->32) against torch fp64."""
    B = 5
    s1, s2 = bf(rnd(B, 32, 15, 23, seed=1)), bf(rnd(B, 64, 8, 12, seed=2))
    w, b = rnd(12, 48, 3, 3, seed=3, scale=(9 * 48) ** -0.5), rnd(12, seed=4)
    cat_ref = torch.cat([F.pixel_shuffle(s2.double(), 2), F.interpolate(s1.double(), size=(16, 24), mode="bilinear", align_corners=True)], dim=1)
    t1 = s1.permute(0, 2, 3, 1).contiguous().to(BF).cuda()
    t2 = s2.permute(0, 2, 3, 1).contiguous().to(BF).cuda()
    cat = tc.shuffle_upsample_cat(t2, 8, 12, t1, 15, 23, 64)
    assert cat.shape == (B, 16, 24, 64) and float(cat[..., 48:].abs().max()) == 0.0
    check_bf16(cat[..., :48].permute(0, 3, 1, 2), cat_ref, "shuffle+upsample cat")
    assert torch.equal(cat[..., :16].permute(0, 3, 1, 2).float().cpu(), F.pixel_shuffle(s2, 2))          # the shuffle is a pure copy
    wp = torch.zeros(32, 64, 3, 3)
    wp[:12, :48] = w
    bp = torch.zeros(32)
    bp[:12] = b
    out = tc.conv3x3_same(cat, tc.pack_conv3x3_weight(wp.cuda()), bp.cuda())
    want = F.conv2d(cat[..., :48].permute(0, 3, 1, 2).double().cpu(), bf(w).double(), b.double(), padding=1)
    check_bf16(out[..., :12].permute(0, 3, 1, 2), want, "down_sample conv (padding=1)")
    assert float(out[..., 12:].abs().max()) == 0.0
    # a larger dense case for the same-padding conv, Cin = Cout = 64, odd extents, ReLU
    x = bf(rnd(3, 64, 21, 35, seed=5))
    w2, b2 = bf(rnd(64, 64, 3, 3, seed=6, scale=(9 * 64) ** -0.5)), rnd(64, seed=7)
    got = tc.conv3x3_same(x.permute(0, 2, 3, 1).contiguous().to(BF).cuda(), tc.pack_conv3x3_weight(w2.cuda()), b2.cuda(), relu=True)
    check_bf16(got.permute(0, 3, 1, 2), F.relu(F.conv2d(x.double(), w2.double(), b2.double(), padding=1)), "same conv 64->64")


@pytest.mark.parametrize("N,H,W", [(2, 40, 52), (1, 260, 346), (3, 21, 13)])
def test_stem_lookup_fused_into_e12(cuda_lib, N, H, W):
    """Binary-mask stem as a table lookup inside the e12 kernel (learner_models.py:489-491,533-535) against PyTorch fp64
    on the bf16-rounded weights, and against the two-kernel path (tensor-core stem, then e12 + pool)."""
    g = torch.Generator().manual_seed(5)
    mask = (torch.rand(N, 1, H, W, generator=g) < 0.35).float()
    w1, b1 = rnd(32, 1, 3, 3, seed=1, scale=0.5), rnd(32, seed=2, scale=0.2)
    w2, b2 = bf(rnd(32, 32, 3, 3, seed=3, scale=(9 * 32) ** -0.5)), rnd(32, seed=4, scale=0.1)
    e11 = F.relu(F.conv2d(mask.double(), bf(w1).double(), b1.double())).to(BF).double()      # the kernel keeps e11 as bf16
    want = F.relu(F.conv2d(e11, w2.double(), b2.double()))
    out, pooled = tc.stem_e12_pool(mask.cuda(), w1.cuda(), b1.cuda(), tc.pack_conv3x3_weight(w2.cuda()), b2.cuda())
    assert (out.vh, out.vw) == (H - 4, W - 4) and out.data.shape == (N, H, W, 32)
    check_bf16(tc.grid_to_nchw(out.data, H - 4, W - 4), want, "stem+e12")
    check_bf16(tc.grid_to_nchw(pooled.data, pooled.vh, pooled.vw), F.max_pool2d(want, 2), "stem+e12 pool")
    ref_out, ref_pool = tc.conv3x3_pool(tc.stem_conv3x3(mask.cuda(), w1.cuda(), b1.cuda()), tc.pack_conv3x3_weight(w2.cuda()), b2.cuda())
    d = (tc.grid_to_nchw(out.data, H - 4, W - 4) - tc.grid_to_nchw(ref_out.data, H - 4, W - 4)).abs()
    assert d.max().item() <= 3e-2 and d.mean().item() <= 1e-3       # e11 values may differ by one bf16 ulp (summation order)


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 24, 40, 32, 32), (1, 72, 152, 32, 32), (1, 19, 11, 64, 64)])
def test_conv3x3_with_fused_1x1_output(cuda_lib, N, H, W, Cin, Cout):
    """unet_d42 + unet_out (learner_models.py:581-583): 3x3 valid conv + ReLU with the 1x1 conv to one channel in the epilogue."""
    x = bf(rnd(N, Cin, H, W, seed=1))
    w = bf(rnd(Cout, Cin, 3, 3, seed=2, scale=(9 * Cin) ** -0.5))
    b = rnd(Cout, seed=3)
    w1, b1 = bf(rnd(Cout, seed=4, scale=Cout ** -0.5)), rnd(1, seed=5)
    act = F.relu(F.conv2d(x.double(), w.double(), b.double())).to(BF).double()          # the epilogue rounds the activation to bf16 first
    want = (act * w1.double().view(1, -1, 1, 1)).sum(1) + b1.double()
    g = tc.nchw_to_grid(x.cuda(), H, W)
    out = tc.conv3x3_out1(g, tc.pack_conv3x3_weight(w.cuda()), b.cuda(), w1.cuda(), b1.cuda())
    got = out[:, :H - 2, :W - 2].cpu().double()
    assert (got - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())


def test_form_patterns_equals_form_input_then_patterns(cuda_lib):
    """evfly_form_patterns = form_input(form_BEV = 2) + evfly_stem_patterns in one pass: same patterns, same in-place cutoff
    of the caller's frames (learner_models.py:477), NaN pixels count as set (F8b)."""
    from evfly_b200 import _lib, ops
    g = torch.Generator().manual_seed(9)
    N, H, W = 3, 41, 53
    fr = (torch.randn(N, 1, H, W, generator=g) * (torch.rand(N, 1, H, W, generator=g) < 0.4)).cuda()
    fr[0, 0, 5, 7] = float("nan")
    fr[1, 0, H - 1, W - 1] = 5e-4              # a corner pixel below the cutoff: owned by the last thread row / segment
    fr[2, 0, H - 2, 3] = -5e-4
    a = fr.clone()
    mask = ops.form_input(a, 2, 1e-3)
    pat_a = torch.empty((N, H - 2, W - 2), dtype=torch.int16, device="cuda")
    _lib.check(cuda_lib.evfly_stem_patterns(_lib.ptr(mask), pat_a.data_ptr(), N, H, W, _lib.stream_ptr()))
    b = fr.clone()
    pat_b = torch.empty_like(pat_a)
    _lib.check(cuda_lib.evfly_form_patterns(_lib.ptr(b), 1e-3, pat_b.data_ptr(), N, H, W, _lib.stream_ptr()))
    assert torch.equal(pat_a, pat_b)
    assert torch.equal(torch.nan_to_num(a, nan=123.0), torch.nan_to_num(b, nan=123.0))
