"""GPU: every L3 operator entry point (through the C ABI) against plain PyTorch on the CPU
(float64 truth, fp32 tolerance rtol 1e-5). Shapes are the ones the reference's layers use plus
ragged/edge cases (tiles that do not divide, strided views, groups, padding)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from evfly_b200 import ops

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-5


def close(got, want, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(got.detach().cpu().double().numpy(), want.double().numpy(), rtol=rtol, atol=atol)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


CONV_CASES = [
    # N, Cin, H,  W,  Cout, k, s, p, groups, act
    (2, 1, 60, 90, 32, 7, 4, 3, 1, None),          # stage-1 patch embed
    (2, 32, 15, 23, 64, 3, 2, 1, 1, None),         # stage-2 patch embed
    (1, 1, 38, 45, 32, 3, 1, 0, 1, "relu"),        # unet_e11-like, K = 9 < one chunk
    (1, 32, 21, 19, 64, 3, 1, 0, 1, "relu"),       # ragged pixel tile
    (1, 70, 9, 11, 130, 3, 1, 0, 1, "relu"),       # ragged K and Cout tiles
    (2, 256, 15, 23, 256, 3, 1, 1, 32, "gelu"),    # MixFFN grouped 3x3, 8 -> 8 per group
    (3, 32, 15, 23, 32, 8, 8, 0, 1, None),         # attention reduction conv (drops a border)
    (2, 48, 16, 24, 12, 3, 1, 1, 1, None),         # down_sample
    (2, 32, 20, 28, 1, 1, 1, 0, 1, None),          # unet_out
    (1, 2, 60, 90, 4, 5, 3, 0, 1, "leaky_relu"),
    (1, 8, 12, 12, 8, 3, 1, 1, 1, "tanh"),
    (1, 8, 12, 12, 8, 3, 1, 1, 1, "sigmoid"),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_vs_torch(cuda_lib, case):
    N, Cin, H, W, Cout, k, s, p, g, act = case
    x, w, b = rnd(N, Cin, H, W, seed=1), rnd(Cout, Cin // g, k, k, seed=2, scale=(Cin // g * k * k) ** -0.5), rnd(Cout, seed=3)
    want = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p, groups=g)
    want = {None: lambda v: v, "relu": F.relu, "gelu": F.gelu, "leaky_relu": F.leaky_relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[act](want)
    got = ops.conv2d(x.cuda(), w.cuda(), b.cuda(), stride=s, pad=p, groups=g, act=act)
    close(got, want)


def test_conv2d_strided_views_post_affine_and_residual(cuda_lib):
    # channel-last input view, channel-last output view, post affine, residual: the ViT glue
    B, H, W, C, Co = 2, 15, 23, 32, 64
    tok = rnd(B, H * W, C, seed=4)
    w, b = rnd(Co, C, 3, 3, seed=5, scale=0.06), rnd(Co, seed=6)
    sc, sh = rnd(Co, seed=7), rnd(Co, seed=8)
    x_bchw = tok.view(B, H, W, C).permute(0, 3, 1, 2)
    res = rnd(B, H * W, Co, seed=9)
    want = F.relu(F.conv2d(x_bchw.double(), w.double(), b.double(), padding=1)) * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
    want = want + res.double().view(B, H, W, Co).permute(0, 3, 1, 2)
    d_tok, d_res = tok.cuda(), res.cuda()
    out = torch.empty((B, H * W, Co), device="cuda")
    ov = out.view(B, H, W, Co).permute(0, 3, 1, 2)
    ops.conv2d(d_tok.view(B, H, W, C).permute(0, 3, 1, 2), w.cuda(), b.cuda(), pad=1, act="relu", post=(sc.cuda(), sh.cuda()),
               res_view=d_res.view(B, H, W, Co).permute(0, 3, 1, 2), out_view=ov)
    close(ov, want)


def test_linear_and_strided_rows(cuda_lib):
    x, w, b = rnd(37, 517, seed=1), rnd(512, 517, seed=2, scale=0.04), rnd(512, seed=3)
    close(ops.linear(x.cuda(), w.cuda(), b.cuda()), F.linear(x.double(), w.double(), b.double()))
    # output and residual into column slices of a wider buffer (metadata concat)
    buf = torch.zeros((37, 600), device="cuda")
    res = rnd(37, 512, seed=4)
    rbuf = torch.zeros((37, 700), device="cuda")
    rbuf[:, 100:612] = res.cuda()
    ops.linear(x.cuda(), w.cuda(), None, act="leaky_relu", res2d=None, out2d=buf[:, 50:562])
    close(buf[:, 50:562], F.leaky_relu(F.linear(x.double(), w.double())))
    assert not buf[:, :50].any() and not buf[:, 562:].any()
    out = ops.linear(x.cuda(), w.cuda(), b.cuda(), res2d=res.cuda())
    close(out, F.linear(x.double(), w.double(), b.double()) + res.double())


def test_conv_transpose_k_equals_stride(cuda_lib):
    from evfly_b200.vitfly_models import conv_transpose, pack_conv_transpose
    for cin, cout, k, H, W in [(512, 256, 2, 8, 13), (8, 4, 3, 20, 30)]:
        m = torch.nn.ConvTranspose2d(cin, cout, kernel_size=k, stride=k)
        x = rnd(2, cin, H, W, seed=3)
        want = m.double()(x.double()).detach()
        m = m.float().cuda()
        cat = torch.zeros((2, cout + 5, k * H, k * W), device="cuda")
        with torch.no_grad():
            conv_transpose(x.cuda(), pack_conv_transpose(m), m.bias, cat[:, 5:])
        close(cat[:, 5:], want)
        assert not cat[:, :5].any()


@pytest.mark.parametrize("k,s,mode,neg", [(2, 2, "max", False), (2, 1, "max", True), (3, 1, "avg", False), (2, 3, "max", False), (3, 1, "max", True), (2, 1, "avg", False)])
def test_pool2d(cuda_lib, k, s, mode, neg):
    x = rnd(2, 5, 25, 35, seed=2)
    if mode == "max":
        want = -F.max_pool2d(-x, k, s) if neg else F.max_pool2d(x, k, s)
        got = ops.pool2d(x.cuda(), k, s, "max", negate_in=neg, negate_out=neg)
    else:
        want, got = F.avg_pool2d(x, k, s), ops.pool2d(x.cuda(), k, s, "avg")
    close(got, want)
    got = ops.pool2d(x.cuda(), k, s, mode, negate_in=True)          # DynamicConvNet: pool(-x)
    close(got, (F.max_pool2d if mode == "max" else F.avg_pool2d)(-x, k, s))


@pytest.mark.parametrize("src,dst,align", [((260, 346), (60, 90), False), ((15, 23), (16, 24), True), ((25, 35), (16, 26), False),
                                            ((256, 342), (72, 152), False), ((68, 148), (260, 346), False), ((7, 5), (7, 5), False), ((4, 4), (1, 1), True)])
def test_resize_bilinear(cuda_lib, src, dst, align):
    x = rnd(2, 3, *src, seed=5)
    want = F.interpolate(x, size=dst, mode="bilinear", align_corners=align)
    close(ops.resize_bilinear(x.cuda(), dst, align_corners=align), want, atol=2e-6)
    # strided input view + output into a concat slice + scale/clip epilogue
    cat = torch.zeros((2, 7, *dst), device="cuda")
    xv = x.cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)       # channel-last storage
    ops.resize_bilinear(xv, dst, align_corners=align, out_view=cat[:, 2:5], mul=2.0, lo=0.0, hi=1.0)
    close(cat[:, 2:5], torch.clip(want * 2, 0, 1), atol=4e-6)
    assert not cat[:, :2].any() and not cat[:, 5:].any()


@pytest.mark.parametrize("C", [32, 64, 100, 1024])
def test_layernorm(cuda_lib, C):
    x, g, b = rnd(77, C, seed=1) * 3 + 1, rnd(C, seed=2), rnd(C, seed=3)
    close(ops.layernorm(x.cuda(), g.cuda(), b.cuda()), F.layer_norm(x.double(), (C,), g.double(), b.double()))


@pytest.mark.parametrize("B,N,C,heads,nkv", [(3, 345, 32, 1, 2), (3, 96, 64, 2, 6), (1, 5, 64, 4, 32)])
def test_attention_small(cuda_lib, B, N, C, heads, nkv):
    q, kv = rnd(B, N, C, seed=1), rnd(B, nkv, 2 * C, seed=2)
    d = C // heads
    kvr = kv.double().reshape(B, nkv, 2, heads, d).permute(2, 0, 3, 1, 4)
    qr = q.double().reshape(B, N, heads, d).permute(0, 2, 1, 3)
    att = torch.softmax(qr @ kvr[0].transpose(-2, -1) / (C / heads) ** 0.5, dim=-1)
    want = (att @ kvr[1]).transpose(1, 2).reshape(B, N, C)
    close(ops.attention_small(q.cuda(), kv.cuda(), heads), want)


def test_map4d_pixel_shuffle_form_input_velpred(cuda_lib):
    x = rnd(2, 6, 9, 7, seed=1)
    close(ops.map4d(x.cuda(), mul=2.0, lo=0.0, hi=1.0), torch.clip(x * 2, 0, 1))
    assert torch.equal(ops.map4d(x.cuda()[:, 1:4, 2:6, 1:5]).cpu(), x[:, 1:4, 2:6, 1:5])          # crop copy is exact
    v = torch.tensor([[4.0], [5.5], [3.3]])
    assert torch.equal(ops.map4d(v.cuda(), div=10.0).cpu(), v / 10) and torch.equal(ops.map4d(v.cuda(), mul=0.1).cpu(), v * 0.1)
    ps = rnd(2, 64, 8, 12, seed=2)
    cat = torch.zeros((2, 48, 16, 24), device="cuda")
    ops.pixel_shuffle(ps.cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2), 2, cat[:, :16])
    assert torch.equal(cat[:, :16].cpu(), F.pixel_shuffle(ps, 2)) and not cat[:, 16:].any()
    fr = (torch.randint(-6, 7, (2, 1, 20, 30)) / 4.0) * (torch.rand(2, 1, 20, 30) < 0.4)
    fr[0, 0, 0, 0] = 5e-4
    fr[0, 0, 0, 1] = float("nan")
    for bev, want in ((0, None), (1, None), (2, None)):
        d = fr.clone().cuda()
        got = ops.form_input(d, bev, 1e-3).cpu()
        ref = fr.clone()
        ref[ref.abs() < 1e-3] = 0.0
        assert torch.equal(d.cpu().nan_to_num(7.0), ref.nan_to_num(7.0))       # caller's tensor is mutated the same way
        if bev == 0:
            pos = torch.where(ref > 0, ref, torch.zeros_like(ref))
            assert torch.equal(got.nan_to_num(7.0), torch.cat([pos, pos], 1).nan_to_num(7.0))
        elif bev == 1:
            assert torch.equal(got.nan_to_num(7.0), ref.abs().nan_to_num(7.0))
        else:
            m = torch.zeros_like(ref)
            m[ref != 0.0] = 1.0
            assert torch.equal(got, m) and got[0, 0, 0, 1] == 1.0          # NaN != 0 -> 1
    y = torch.tanh(rnd(5, 1, seed=3))
    close(ops.velpred_unit(y.cuda()), torch.cat([torch.sqrt(1 - y ** 2), y, torch.zeros_like(y)], 1))


@pytest.mark.parametrize("T,inp,H,layers,bias", [(7, 517, 128, 3, True), (5, 665, 395, 2, False), (1, 40, 16, 1, True)])
def test_lstm_sequence_vs_nn_lstm(cuda_lib, T, inp, H, layers, bias):
    from evfly_b200._modbase import pack_lstm, run_lstm
    torch.manual_seed(0)
    m = torch.nn.LSTM(input_size=inp, hidden_size=H, num_layers=layers, bias=bias).eval()
    x = rnd(T, inp, seed=1)
    h0, c0 = rnd(layers, H, seed=2) * 0.3, rnd(layers, H, seed=3) * 0.3
    with torch.no_grad():
        want, (hn, cn) = m.double()(x.double(), (h0.double(), c0.double()))
        want0, _ = m(x.double())
        mc = m.float().cuda()
        got, (h, c) = run_lstm(ops, pack_lstm(mc), x.cuda(), (h0, c0), H)
        got0, _ = run_lstm(ops, pack_lstm(mc), x.cuda(), None, H)
    close(got, want); close(h, hn); close(c, cn); close(got0, want0)


def test_convlstm_vs_reference_formula(cuda_lib):
    from evfly_b200.ConvLSTM_pytorch.convlstm import ConvLSTM
    from oracle.model_oracle import convlstm_seq
    torch.manual_seed(1)
    m = ConvLSTM(input_dim=64, hidden_dim=[64], num_layers=1, kernel_size=(1, 1), bias=False, batch_first=True).eval()
    x = rnd(4, 64, 8, 13, seed=2)
    sd = {"lstm." + k: v.detach() for k, v in m.state_dict().items()}
    with torch.no_grad():
        want, st = convlstm_seq(sd, "lstm", x, None, hidden=64)
        want2, _ = convlstm_seq(sd, "lstm", x.flip(0), st, hidden=64)
        m = m.cuda()
        out, state = m(x.unsqueeze(0).cuda(), None)
        out2, _ = m(x.flip(0).unsqueeze(0).cuda(), state)
    close(out[0][0], want); close(state[0][0], st[0][0]); close(state[0][1], st[0][1]); close(out2[0][0], want2)
    # 3x3 kernel with bias, two layers, batch 2, time-major: the general constructor path
    torch.manual_seed(2)
    m = ConvLSTM(input_dim=5, hidden_dim=[6, 4], num_layers=2, kernel_size=[(3, 3), (3, 3)], bias=True, batch_first=False, return_all_layers=True).eval()
    x = rnd(3, 2, 5, 7, 9, seed=4)
    with torch.no_grad():
        cur = x.permute(1, 0, 2, 3, 4)
        for li, hid in enumerate((6, 4)):
            sdl = {"l.cell_list.0.conv.weight": m.state_dict()[f"cell_list.{li}.conv.weight"], "l.cell_list.0.conv.bias": m.state_dict()[f"cell_list.{li}.conv.bias"]}
            cur = torch.stack([convlstm_seq(sdl, "l", cur[b], None, hidden=hid)[0] for b in range(2)])
        outs, _ = m.cuda()(x.cuda())
    close(outs[1], cur)
