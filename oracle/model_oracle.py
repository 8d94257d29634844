"""Functional (state_dict-in, tensors-out) fp32 PyTorch-CPU restatement of the reference's model
forward for the hot path. TEST INFRASTRUCTURE: the checker for the CUDA kernels and the CPU
baseline of bench.py; never imported by evfly_b200/.

It is a floating-point path, so a torch fp32 reference is the oracle (tolerances in the tests:
rtol 1e-5 fp32 path, 1e-2 bf16 path -- BASELINE.json north_star). Pinned against the reference:
tests/golden/make_golden_models.py runs the reference's own nn.Modules (imported from
/root/reference/learner) on seeded synthetic checkpoints and inputs and commits their outputs;
tests/test_oracle_models.py checks this file against those vectors.

Follows (paths relative to the evfly tree):
  learner/ViTsubmodules.py:15-148          patch merge, efficient self-attention, MixFFN, stage
  learner/vitfly_models.py:18-263          refine_inputs, ConvNet, LSTMNet, LSTMNetVIT, ViT, UNetConvLSTMNet
  learner/ConvLSTM_pytorch/convlstm.py:38-53,136-176   ConvLSTM cell and sequence loop
  learner/learner_models.py:476-636        OrigUNet (form_input, encoder, ConvLSTM, decoder,
                                           velpred head) and OrigUNet_w_VITFLY_ViTLSTM
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------
def sn_weight(sd, prefix):
    """Old-style torch.nn.utils.spectral_norm in eval mode: W / sigma, sigma = u^T (W v); no
    power iteration (vitfly_models.py:123,126 wrap nn.Linear)."""
    w = sd[prefix + ".weight_orig"]
    u, v = sd[prefix + ".weight_u"], sd[prefix + ".weight_v"]
    sigma = torch.dot(u, torch.mv(w.reshape(w.shape[0], -1), v))
    return w / sigma


def linear_sn(sd, prefix, x):
    return F.linear(x, sn_weight(sd, prefix), sd[prefix + ".bias"])


def lstm_seq(sd, prefix, x, state, hidden, layers, bias=True):
    """nn.LSTM on an UNBATCHED sequence x [T, in] (the reference feeds (N,feat) 2-D tensors, so
    N is time -- SURVEY.md F2). Gate order i,f,g,o. state = (h0,c0) each [layers, hidden] or None."""
    T = x.shape[0]
    if state is None:
        h0 = x.new_zeros(layers, hidden)
        c0 = x.new_zeros(layers, hidden)
    else:
        h0, c0 = state
    inp = x
    hT, cT = [], []
    for l in range(layers):
        w_ih, w_hh = sd[f"{prefix}.weight_ih_l{l}"], sd[f"{prefix}.weight_hh_l{l}"]
        b = 0.0
        if bias:
            b = sd[f"{prefix}.bias_ih_l{l}"] + sd[f"{prefix}.bias_hh_l{l}"]
        gx = inp @ w_ih.t() + b
        h, c = h0[l], c0[l]
        outs = []
        for t in range(T):
            g = gx[t] + w_hh @ h
            i, f, gg, o = g.chunk(4)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs)
        hT.append(h)
        cT.append(c)
    return inp, (torch.stack(hT), torch.stack(cT))


# ---------------------------------------------------------------------------------------------
# ViTsubmodules.py
# ---------------------------------------------------------------------------------------------
def overlap_patch_merging(sd, p, x, stride, padding):
    y = F.conv2d(x, sd[p + ".cn1.weight"], sd[p + ".cn1.bias"], stride=stride, padding=padding)
    B, C, H, W = y.shape
    tok = y.flatten(2).transpose(1, 2)
    tok = F.layer_norm(tok, (C,), sd[p + ".layerNorm.weight"], sd[p + ".layerNorm.bias"])
    return tok, H, W


def efficient_self_attention(sd, p, x, H, W, reduction, heads):
    B, N, C = x.shape
    d = C // heads
    red = F.conv2d(x.permute(0, 2, 1).reshape(B, C, H, W), sd[p + ".cn1.weight"], sd[p + ".cn1.bias"], stride=reduction)
    red = red.reshape(B, C, -1).permute(0, 2, 1)
    red = F.layer_norm(red, (C,), sd[p + ".ln1.weight"], sd[p + ".ln1.bias"])
    kv = F.linear(red, sd[p + ".keyValueExtractor.weight"], sd[p + ".keyValueExtractor.bias"])
    kv = kv.reshape(B, -1, 2, heads, d).permute(2, 0, 3, 1, 4)          # [kv][B][head][n_kv][d]
    k, v = kv[0], kv[1]
    q = F.linear(x, sd[p + ".query.weight"], sd[p + ".query.bias"]).reshape(B, N, heads, d).permute(0, 2, 1, 3)
    att = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(C / heads), dim=-1)
    out = (att @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(out, sd[p + ".finalLayer.weight"], sd[p + ".finalLayer.bias"])


def mix_ffn(sd, p, x, H, W, channels):
    y = F.linear(x, sd[p + ".mlp1.weight"], sd[p + ".mlp1.bias"])
    B, N, Ce = y.shape
    y = y.transpose(1, 2).reshape(B, Ce, H, W)
    y = F.conv2d(y, sd[p + ".depthwise.weight"], sd[p + ".depthwise.bias"], padding=1, groups=channels)
    y = F.gelu(y.flatten(2).transpose(1, 2))
    return F.linear(y, sd[p + ".mlp2.weight"], sd[p + ".mlp2.bias"])


def mix_transformer_stage(sd, p, x, *, channels, stride, padding, n_layers, reduction, heads):
    B = x.shape[0]
    tok, H, W = overlap_patch_merging(sd, p + ".patchMerge", x, stride, padding)
    for i in range(n_layers):
        tok = tok + efficient_self_attention(sd, f"{p}._attn.{i}", tok, H, W, reduction, heads)
        tok = tok + mix_ffn(sd, f"{p}._ffn.{i}", tok, H, W, channels)
        tok = F.layer_norm(tok, (channels,), sd[f"{p}._lNorm.{i}.weight"], sd[f"{p}._lNorm.{i}.bias"])
    return tok.reshape(B, H, W, channels).permute(0, 3, 1, 2)


STAGE1 = dict(channels=32, stride=4, padding=3, n_layers=2, reduction=8, heads=1)
STAGE2 = dict(channels=64, stride=2, padding=1, n_layers=2, reduction=4, heads=2)


# ---------------------------------------------------------------------------------------------
# vitfly_models.py
# ---------------------------------------------------------------------------------------------
def refine(depth, quat):
    if quat is None:
        quat = depth.new_zeros(depth.shape[0], 4)
        quat[:, 0] = 1
    if depth.shape[-2] != 60 or depth.shape[-1] != 90:
        depth = F.interpolate(depth, size=(60, 90), mode="bilinear")
    return depth, quat


def vit_encoder_features(sd, depth):
    """shared by LSTMNetVIT and ViT: two stages -> cat[pixel_shuffle(s2), upsample(s1)] ->
    conv 48->12 -> flatten 4608 (vitfly_models.py:136-143 / :174-180)."""
    s1 = mix_transformer_stage(sd, "encoder_blocks.0", depth, **STAGE1)
    s2 = mix_transformer_stage(sd, "encoder_blocks.1", s1, **STAGE2)
    cat = torch.cat([F.pixel_shuffle(s2, 2),
                     F.interpolate(s1, size=(16, 24), mode="bilinear", align_corners=True)], dim=1)
    return F.conv2d(cat, sd["down_sample.weight"], sd["down_sample.bias"], padding=1).flatten(1)


def lstmnet_vit(sd, depth, desvel, quat=None, state=None):
    """LSTMNetVIT.forward -> (vel [N,3], (h,c) each [3,128])."""
    depth, quat = refine(depth, quat)
    feat = linear_sn(sd, "decoder", vit_encoder_features(sd, depth))
    seq = torch.cat([feat, desvel / 10, quat], dim=1).float()
    out, hc = lstm_seq(sd, "lstm", seq, state, hidden=128, layers=3)
    return linear_sn(sd, "nn_fc2", out), hc


def vit(sd, depth, desvel, quat=None):
    """ViT.forward -> vel [N,3] (stateless)."""
    depth, quat = refine(depth, quat)
    feat = F.linear(vit_encoder_features(sd, depth), sd["decoder.weight"], sd["decoder.bias"])
    x = torch.cat([feat, desvel / 10, quat], dim=1).float()
    x = F.leaky_relu(linear_sn(sd, "nn_fc1", x))
    return linear_sn(sd, "nn_fc2", x)


def _bn(sd, p, x, eps=1e-5):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, eps)


def convnet(sd, depth, desvel, quat=None):
    """ConvNet.forward (vitfly_models.py:51-70)."""
    depth, quat = refine(depth, quat)
    x = F.relu(F.conv2d(depth, sd["conv1.weight"], sd["conv1.bias"], stride=3))
    x = -F.max_pool2d(-_bn(sd, "bn1", x), 2, 1)
    x = F.avg_pool2d(F.relu(F.conv2d(x, sd["conv2.weight"], sd["conv2.bias"], stride=2)), 3, 1)
    x = torch.cat([x.flatten(1), desvel * 0.1, quat], dim=1).float()
    x = F.leaky_relu(F.linear(x, sd["fc0.weight"]))
    x = F.leaky_relu(F.linear(x, sd["fc1.weight"]))
    x = torch.tanh(F.linear(x, sd["fc2.weight"]))
    return F.linear(x, sd["fc3.weight"], sd["fc3.bias"])


def lstmnet(sd, depth, desvel, quat=None, state=None):
    """LSTMNet.forward (vitfly_models.py:92-109)."""
    depth, quat = refine(depth, quat)
    x = F.relu(F.conv2d(depth, sd["conv1.weight"], sd["conv1.bias"], stride=3, padding=1))
    x = -F.max_pool2d(-_bn(sd, "bn1", x), 3, 1)
    x = F.relu(F.conv2d(x, sd["conv2.weight"], sd["conv2.bias"], stride=2))
    x = F.avg_pool2d(_bn(sd, "bn2", x), 3, 1)
    seq = torch.cat([x.flatten(1), desvel * 0.1, quat], dim=1).float()
    out, hc = lstm_seq(sd, "lstm", seq, state, hidden=395, layers=2, bias=False)
    out = F.leaky_relu(linear_sn(sd, "fc1", out))
    out = F.leaky_relu(linear_sn(sd, "fc2", out))
    return linear_sn(sd, "fc3", out), hc


def unet_convlstm_net(sd, depth, desvel, quat=None, state=None):
    """UNetConvLSTMNet.forward (vitfly_models.py:231-263)."""
    img, quat = refine(depth, quat)
    c = lambda n, x, **kw: F.conv2d(x, sd[n + ".weight"], sd[n + ".bias"], **kw)
    e1 = F.relu(c("unet_e12", F.relu(c("unet_e11", img, padding=1)), padding=1))
    e2 = F.relu(c("unet_e22", F.relu(c("unet_e21", F.max_pool2d(e1, 2, 3), padding=1)), padding=1))
    e3 = F.relu(c("unet_e32", F.relu(c("unet_e31", F.max_pool2d(e2, 2, 2), padding=1)), padding=1))
    up1 = F.conv_transpose2d(e3, sd["unet_upconv1.weight"], sd["unet_upconv1.bias"], stride=2)
    d1 = F.relu(c("unet_d12", F.relu(c("unet_d11", torch.cat([up1, e2], 1), padding=1)), padding=1))
    up2 = F.conv_transpose2d(d1, sd["unet_upconv2.weight"], sd["unet_upconv2.bias"], stride=3)
    d2 = F.relu(c("unet_d22", F.relu(c("unet_d21", torch.cat([up2, e1], 1), padding=1)), padding=1))
    y_unet = c("unet_out", d2)
    xc = torch.cat([img, y_unet], 1)
    yc = -F.max_pool2d(-F.relu(_bn(sd, "conv_bn1", c("conv_conv1", xc, stride=3))), 2, 1)
    yc = F.avg_pool2d(F.relu(c("conv_conv2", yc, stride=2)), 2, 1)
    seq = torch.cat([yc.flatten(1), e3.flatten(1), desvel * 0.1, quat], dim=1).float()
    out, hc = lstm_seq(sd, "lstm", seq, state, hidden=200, layers=2, bias=False)
    out = F.leaky_relu(linear_sn(sd, "nn_fc1", out))
    out = F.leaky_relu(linear_sn(sd, "nn_fc2", out))
    return linear_sn(sd, "nn_fc3", out), hc


# ---------------------------------------------------------------------------------------------
# ConvLSTM_pytorch/convlstm.py (1 layer, 1x1 kernel, no bias as OrigUNet builds it)
# ---------------------------------------------------------------------------------------------
def convlstm_seq(sd, prefix, x_seq, state, hidden=512):
    """x_seq [T,C,h,w] = the reference's [1,T,C,h,w]; gate order i,f,o,g (convlstm.py:44).
    Returns (h_seq [T,hidden,h,w], [[h,c]] with h,c [1,hidden,h,w])."""
    w = sd[prefix + ".cell_list.0.conv.weight"]
    b = sd.get(prefix + ".cell_list.0.conv.bias")
    T, _, hh, ww = x_seq.shape
    if state is None:
        h = x_seq.new_zeros(1, hidden, hh, ww)
        c = x_seq.new_zeros(1, hidden, hh, ww)
    else:
        h, c = state[0]
    pad = w.shape[-1] // 2
    outs = []
    for t in range(T):
        g = F.conv2d(torch.cat([x_seq[t:t + 1], h], dim=1), w, b, padding=pad)
        gi, gf, go, gg = torch.split(g, hidden, dim=1)
        c = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
        h = torch.sigmoid(go) * torch.tanh(c)
        outs.append(h[0])
    return torch.stack(outs), [[h, c]]


# ---------------------------------------------------------------------------------------------
# learner_models.py: OrigUNet
# ---------------------------------------------------------------------------------------------
def form_input(x, form_bev, cutoff):
    """learner_models.py:476-494 (the in-place zeroing of the caller's tensor is reproduced)."""
    x[x.abs() < cutoff] = 0.0
    if form_bev == 0:
        # The reference builds the 2-channel tensor as zeros_like(x).expand(-1, 2, -1, -1): a
        # stride-0 view whose two channels ALIAS one buffer, so the second assignment (positive
        # part) overwrites the first (|negative| part). Its actual output is therefore
        # [pos, pos]; reproduced here because the drop-in must return what the reference returns.
        pos = torch.where(x > 0, x, torch.zeros_like(x))
        return torch.cat([pos, pos], dim=1)
    if form_bev == 1:
        return x.abs()
    if form_bev == 2:
        return (x != 0.0).to(x.dtype)
    raise ValueError(form_bev)


SKIPS = [((25, 35), (16, 26)), ((58, 79), (24, 44)), ((124, 167), (40, 80)), ((256, 342), (72, 152))]


def _skip(y, big, small, skip_type):
    if skip_type == "crop":
        return y[:, :, big[0] // 2 - small[0] // 2: big[0] // 2 + small[0] // 2,
                 big[1] // 2 - small[1] // 2: big[1] // 2 + small[1] // 2]
    if skip_type == "interp":
        return F.interpolate(y, size=small, mode="bilinear", align_corners=False)
    if skip_type == "none":
        return None
    raise ValueError(skip_type)


def velpred11_head(sd, y_upconv, enc):
    """velpred == 11: DynamicConvNet (conv no-bias -> BN -> act -> [-]pool[-]) x L on y_upconv,
    flatten, DynamicFCNet, VelPredictor num_out=1 -> [sqrt(1-y^2), y, 0]
    (learner_models.py:18-100,102-145,274-336,594-614)."""
    x = y_upconv
    for i in range(enc["num_layers"]):
        x = F.conv2d(x, sd[f"convnet_velpred.layers.conv2d_{i}.weight"], None, stride=enc["kernel_strides"][i])
        x = _bn(sd, f"convnet_velpred.layers.batchnorm_{i}", x)
        act = enc["activations"][i]
        x = {"relu": F.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh, "leaky_relu": F.leaky_relu, "none": lambda v: v}[act](x)
        # add_module(f'invert_{i}') is called twice with the same name: the second registration
        # REPLACES the first, so nn.Sequential runs exactly one InvertLayer, placed where the first
        # was registered (before the pool): x -> pool(-x)   (learner_models.py:76-93)
        if enc["invert_pool_inputs"]:
            x = -x
        k, s = enc["pool_kernels"][i], enc["pool_strides"][i]
        if enc["pool_type"] == "max":
            x = F.max_pool2d(x, k, s)
        elif enc["pool_type"] == "avg":
            x = F.avg_pool2d(x, k, s)
    x = x.flatten(1)
    return x


def fc_head(sd, x, fc):
    acts = {"relu": F.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh, "leaky_relu": F.leaky_relu}
    for i in range(fc["num_layers"]):
        x = acts[fc["activations"][i]](F.linear(x, sd[f"velpred_head.fcnet.layers.fc_{i}.weight"], sd[f"velpred_head.fcnet.layers.fc_{i}.bias"]))
    rad = 1.0 - x.pow(2)
    first = torch.sqrt(torch.clip(rad, 0.0, 1.0)) if (rad < 0).any() else torch.sqrt(rad)
    return torch.cat([first, x, torch.zeros_like(x)], dim=1)


def orig_unet(sd, frames, state=None, *, form_bev=0, num_in_channels=2, cutoff=1e-3, skip_type="crop",
              num_recurrent=(0, 0), velpred=0, enc_params=None, fc_params=None, input_hw=(260, 346)):
    """OrigUNet.forward -> (vel [N,3], (y_interp, y_upconv, (h_unet, h_velpred)))."""
    c = lambda n, x: F.conv2d(x, sd[n + ".weight"], sd[n + ".bias"])
    im = frames
    if num_in_channels == 2 or form_bev > 0:
        im = form_input(im, form_bev, cutoff)
    st = (None, None) if state is None else state
    e1 = F.relu(c("unet_e12", F.relu(c("unet_e11", im))))
    e2 = F.relu(c("unet_e22", F.relu(c("unet_e21", F.max_pool2d(e1, 2)))))
    e3 = F.relu(c("unet_e32", F.relu(c("unet_e31", F.max_pool2d(e2, 2)))))
    e4 = F.relu(c("unet_e42", F.relu(c("unet_e41", F.max_pool2d(e3, 2)))))
    e5 = F.relu(c("unet_e52", F.relu(c("unet_e51", F.max_pool2d(e4, 2)))))
    h_unet = None
    if num_recurrent[0] > 0:
        e5, h_unet = convlstm_seq(sd, "lstm", e5, st[0])
    y = e5
    for lvl, enc in enumerate((e4, e3, e2, e1), start=1):
        up = F.conv_transpose2d(y, sd[f"unet_upconv{lvl}.weight"], sd[f"unet_upconv{lvl}.bias"], stride=2)
        sk = _skip(enc, *SKIPS[lvl - 1], skip_type)
        cat = up if sk is None else torch.cat([sk, up], dim=1)
        y = F.relu(c(f"unet_d{lvl}2", F.relu(c(f"unet_d{lvl}1", cat))))
    y_upconv = c("unet_out", y)
    y_interp = F.interpolate(y_upconv, size=tuple(input_hw), mode="bilinear", align_corners=False)
    vel = torch.tensor([1.0, 0.0, 0.0]).repeat(frames.shape[0], 1)
    if velpred == 11:
        vel = fc_head(sd, velpred11_head(sd, y_upconv, enc_params), fc_params)
    elif velpred != 0:
        raise NotImplementedError("oracle covers velpred 0 and 11 (the shipped configurations)")
    return vel, (y_interp, y_upconv, (h_unet, None))


def split_prefix(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def orig_unet_w_vitlstm(sd, frames, desvel, unet_state=None, vit_state=None, **unet_cfg):
    """OrigUNet_w_VITFLY_ViTLSTM.forward (learner_models.py:629-636)
    -> (vel, (depth, y_upconv, ((h_unet, None), (h,c))))."""
    unet_cfg = dict(unet_cfg)
    unet_cfg["velpred"] = unet_cfg.get("velpred", 0)
    _, (depth, y_upconv, (h_unet, h_vp)) = orig_unet(split_prefix(sd, "origunet."), frames,
                                                    (unet_state, None), **unet_cfg)
    depth_in = torch.clip(depth * 2, 0.0, 1.0)
    vel, hc = lstmnet_vit(split_prefix(sd, "vitfly_vitlstm."), depth_in, desvel, None, vit_state)
    return vel, (depth, y_upconv, ((h_unet, h_vp), hc))


DEPLOYED_UNET_CFG = dict(form_bev=2, num_in_channels=2, cutoff=1e-3, skip_type="interp", num_recurrent=(1, 0), velpred=0)
