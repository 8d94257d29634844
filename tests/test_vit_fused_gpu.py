"""GPU: the fused MixFFN kernel (csrc/vit_fused.cu: mlp1 -> grouped 3x3 conv -> GELU -> mlp2 -> + x -> LayerNorm, one CTA
per sample, tcgen05 with the padded token grid resident in shared memory) against PyTorch fp64 on the same bf16-rounded
tokens and weights (learner/ViTsubmodules.py:85-120,143-146), and against the per-op bf16 path it replaces."""
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


def reference(x, w1, b1, wd, bd, w2, b2, g, beta, H, W, C, round_intermediates):
    """fp64 MixFFN + residual + LayerNorm; round_intermediates: round y1 / y2 to bf16 where the kernel does."""
    B = x.shape[0]
    r = (lambda t: t.to(BF).double()) if round_intermediates else (lambda t: t)
    y = r(x.double() @ w1.double().t() + b1.double())
    y = y.transpose(1, 2).reshape(B, 8 * C, H, W)
    y = F.conv2d(y, wd.double(), bd.double(), padding=1, groups=C)
    y = r(F.gelu(y.flatten(2).transpose(1, 2)))
    y = y @ w2.double().t() + b2.double() + x.double()
    return F.layer_norm(y, (C,), g.double(), beta.double(), 1e-5)


@pytest.mark.parametrize("H,W,C", [(15, 23, 32), (8, 12, 64)])
@pytest.mark.parametrize("B", [1, 3, 40, 700])
def test_fused_mixffn_block(cuda_lib, H, W, C, B):
    Ce = 8 * C
    x = bf(rnd(B, H * W, C, seed=1))
    w1, b1 = bf(rnd(Ce, C, seed=2, scale=C ** -0.5)), rnd(Ce, seed=3, scale=0.1)
    wd, bd = bf(rnd(Ce, 8, 3, 3, seed=4, scale=72 ** -0.5)), rnd(Ce, seed=5, scale=0.1)
    w2, b2 = bf(rnd(C, Ce, seed=6, scale=Ce ** -0.5)), rnd(C, seed=7, scale=0.1)
    g, beta = 1 + 0.1 * rnd(C, seed=8), 0.1 * rnd(C, seed=9)
    img, fb = tc.pack_vit_ffn(*(t.cuda() for t in (w1, b1, wd, bd, w2, b2, g, beta)))
    got = tc.vit_ffn(x.to(BF).cuda(), img, fb, B, H, W, 1e-5).float().cpu().double()
    want = reference(x, w1, b1, wd, bd, w2, b2, g, beta, H, W, C, True)
    err = (got - want).abs()
    # LayerNorm output is O(1): bf16 output rounding (2^-8 relative) + the bf16 rounding of the two intermediates
    assert err.max() <= 4e-2, f"max err {err.max():.4g}"
    assert err.mean() <= 4e-3, f"mean err {err.mean():.4g}"
    exact = reference(x, w1, b1, wd, bd, w2, b2, g, beta, H, W, C, False)
    assert (got - exact).norm() / exact.norm() <= 1e-2


@pytest.mark.parametrize("C,heads,N,n_kv", [(32, 1, 345, 2), (64, 2, 96, 6)])
@pytest.mark.parametrize("B", [1, 5, 400])
def test_fused_attention_block(cuda_lib, C, heads, N, n_kv, B):
    """out = x + finalLayer(softmax(query(x) K^T / sqrt(d)) V) (ViTsubmodules.py:74-83,144) for given K, V."""
    d = C // heads
    x = bf(rnd(B, N, C, seed=1))
    kv = bf(rnd(B, n_kv, 2 * C, seed=2))
    wq, bq = bf(rnd(C, C, seed=3, scale=C ** -0.5)), rnd(C, seed=4, scale=0.1)
    wf, bfin = bf(rnd(C, C, seed=5, scale=C ** -0.5)), rnd(C, seed=6, scale=0.1)
    img, bias = tc.pack_vit_attn(wq.cuda(), bq.cuda(), wf.cuda(), bfin.cuda())
    got = tc.vit_attn(x.to(BF).cuda(), kv.to(BF).cuda(), img, bias, heads).float().cpu().double()
    q = (x.double() @ wq.double().t() + bq.double()).reshape(B, N, heads, d).permute(0, 2, 1, 3)
    kvr = kv.double().reshape(B, n_kv, 2, heads, d).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ kvr[0].transpose(-2, -1) / d ** 0.5, dim=-1) @ kvr[1]
    att = att.transpose(1, 2).reshape(B, N, C).to(BF).double()                      # the kernel rounds the attention row to bf16
    want = x.double() + att @ wf.double().t() + bfin.double()
    err = (got - want).abs()
    assert (err <= 2 ** -7 * want.abs() + 2e-2).all(), f"max err {err.max():.4g}"
    assert err.mean() <= 2 ** -8 * want.abs().mean() + 1e-3


def test_fused_path_equals_per_op_path_in_the_stage(cuda_lib):
    """A whole stage through encode_bf16 with the fused block and with the per-op launches it replaces."""
    import json, os
    import evfly_b200
    from oracle.synth_ckpt import synth_state_dict, synthetic_depth
    from tests.test_models_cpu import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    man = json.load(open(os.path.join(root, "tests", "golden", "state_dict_manifest.json")))
    m = build("LSTMNetVIT")
    m.load_state_dict(synth_state_dict(man["LSTMNetVIT"], 11), strict=True)
    m = evfly_b200.set_precision(m.cuda().eval().float(), "bf16")
    depth = synthetic_depth(3, 24).cuda()
    with torch.no_grad():
        t1, H1, W1 = m.encoder_blocks[0].encode_bf16(depth, True, 24, 60, 90)
        t2, _, _ = m.encoder_blocks[1].encode_bf16(t1, False, 24, H1, W1)
        old = tc.FUSED_FFN_MIN_BATCH
        tc.FUSED_FFN_MIN_BATCH, tc.FUSED_ATTN = 10 ** 9, False
        try:
            u1, _, _ = m.encoder_blocks[0].encode_bf16(depth, True, 24, 60, 90)
            u2, _, _ = m.encoder_blocks[1].encode_bf16(u1, False, 24, H1, W1)
        finally:
            tc.FUSED_FFN_MIN_BATCH, tc.FUSED_ATTN = old, True
    for a, b_, name in ((t1, u1, "stage 1"), (t2, u2, "stage 2")):
        d = (a.float() - b_.float()).abs()
        assert d.max().item() <= 0.15 and d.mean().item() <= 1e-2, (name, d.max().item(), d.mean().item())
