"""Drop-in for evfly's learner/vitfly_models.py: ConvNet, LSTMNet, LSTMNetVIT, ViT,
UNetConvLSTMNet with the reference's constructors (no arguments), state_dict keys and
forward(X) signature, X = [depth [N,1,h,w], desvel [N,1], quat [N,4] or None, (h,c) (optional)].
Returns (vel [N,3], (h,c) or None). N is TIME for the LSTM variants.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.utils.spectral_norm as spectral_norm
from torch.nn import LSTM

from . import ops, tc
from ._modbase import PackedModule, bn_affine, pack_lstm, run_lstm, sn_effective_weight, to_dev
from .ViTsubmodules import *  # noqa: F401,F403  (the reference does the same)
from .ViTsubmodules import MixTransformerEncoderLayer


def refine_inputs(X):
    """vitfly_models.py:18-31: default quaternion [1,0,0,0]; bilinear resize to 60x90.
    Mutates the caller's list like the reference."""
    if X[2] is None:
        X[2] = torch.zeros((X[0].shape[0], 4), dtype=torch.float32, device=X[0].device)
        X[2][:, 0] = 1
    if X[0].shape[-2] != 60 or X[0].shape[-1] != 90:
        X[0] = ops.resize_bilinear(X[0], (60, 90), align_corners=False)
    return X


def _inputs(module, X):
    """Move the caller's tensors next to the weights (plumbing) and apply refine_inputs."""
    dev = module._device()
    module._check_inference()
    X[0] = to_dev(X[0], dev)
    X[1] = to_dev(X[1], dev)
    X[2] = to_dev(X[2], dev)
    return refine_inputs(X)


def _meta_concat(feat_cols: int, pieces, n: int, dev):
    """seq = cat([features, desvel-term, quat], dim=1) without materialising the pieces:
    returns the [n, total] buffer and the view the feature producer writes into."""
    total = feat_cols + sum(p[0].shape[1] for p in pieces)
    seq = torch.empty((n, total), dtype=torch.float32, device=dev)
    col = feat_cols
    for src, mul, div in pieces:
        w = src.shape[1]
        ops.map4d(src, seq[:, col:col + w], mul=mul, div=div)
        col += w
    return seq, seq[:, :feat_cols]


class _ViTEncoder(PackedModule):
    """The part LSTMNetVIT and ViT share (vitfly_models.py:118-130,136-142 / :159-169,174-179)."""

    def _build_encoder(self):
        self.encoder_blocks = nn.ModuleList([
            MixTransformerEncoderLayer(1, 32, patch_size=7, stride=4, padding=3, n_layers=2, reduction_ratio=8, num_heads=1, expansion_factor=8),
            MixTransformerEncoderLayer(32, 64, patch_size=3, stride=2, padding=1, n_layers=2, reduction_ratio=4, num_heads=2, expansion_factor=8)
        ])

    def _build_tail(self):
        self.up_sample = nn.Upsample(size=(16, 24), mode='bilinear', align_corners=True)
        self.pxShuffle = nn.PixelShuffle(upscale_factor=2)
        self.down_sample = nn.Conv2d(48, 12, 3, padding=1)

    precision = 'fp32'

    def _encode(self, depth):
        """[N,1,60,90] -> [N,4608] features."""
        N = depth.shape[0]
        if self.precision == 'bf16':
            b0, b1 = self.encoder_blocks
            t1, H1, W1 = b0.encode_bf16(depth.contiguous(), True, N, depth.shape[2], depth.shape[3])
            t2, H2, W2 = b1.encode_bf16(t1, False, N, H1, W1)
            s1 = tc.grid_to_nchw(t1.view(N, H1, W1, -1), H1, W1)      # [N,32,15,23] fp32 for the small fp32 tail
            s2 = tc.grid_to_nchw(t2.view(N, H2, W2, -1), H2, W2)      # [N,64,8,12]
        else:
            s1 = self.encoder_blocks[0](depth)          # [N,32,15,23] view
            s2 = self.encoder_blocks[1](s1)             # [N,64,8,12] view
        cat = torch.empty((N, 48, 16, 24), dtype=torch.float32, device=depth.device)
        ops.pixel_shuffle(s2, 2, cat[:, :16])
        ops.resize_bilinear(s1, (16, 24), align_corners=True, out_view=cat[:, 16:])
        ds = self.down_sample
        return ops.conv2d(cat, ds.weight, ds.bias, pad=1).view(N, -1)


    TAIL_CIN, TAIL_COUT = 64, 32      # down_sample's 48 -> 12 channels zero-padded to what the halo conv kernel takes

    def _pack_tail_tc(self, decoder_weight):
        """Weights of the tensor-core tail: down_sample as [32, 9*64] bf16 and the decoder Linear re-indexed from the
        reference's NCHW flatten (c*384 + y*24 + x, vitfly_models.py:142) to the NHWC-padded order (y*24 + x)*32 + c."""
        ds = self.down_sample
        w = torch.zeros((self.TAIL_COUT, self.TAIL_CIN, 3, 3), dtype=torch.float32, device=ds.weight.device)
        w[:12, :48] = ds.weight
        b = torch.zeros((self.TAIL_COUT,), dtype=torch.float32, device=ds.weight.device)
        b[:12] = ds.bias
        n_out = decoder_weight.shape[0]
        dec = torch.zeros((n_out, 16, 24, self.TAIL_COUT), dtype=torch.float32, device=decoder_weight.device)
        dec[..., :12] = decoder_weight.view(n_out, 12, 16, 24).permute(0, 2, 3, 1)
        return {"ds_w": tc.pack_conv3x3_weight(w), "ds_b": b, "dec": dec.view(n_out, -1).to(tc.BF16).contiguous()}

    def _encode_tc(self, depth, tail):
        """bf16 path for batches: [N,1,60,90] -> bf16 features [N, 16*24*32] in NHWC-padded order (see _pack_tail_tc);
        everything after the token stages stays bf16 NHWC and on the tensor cores."""
        N = depth.shape[0]
        b0, b1 = self.encoder_blocks
        t1, H1, W1 = b0.encode_bf16(depth.contiguous(), True, N, depth.shape[2], depth.shape[3])
        t2, H2, W2 = b1.encode_bf16(t1, False, N, H1, W1)
        cat = tc.shuffle_upsample_cat(t2, H2, W2, t1, H1, W1, self.TAIL_CIN)          # [N,16,24,64]
        return tc.conv3x3_same(cat, tail["ds_w"], tail["ds_b"]).view(N, -1)


class LSTMNetVIT(_ViTEncoder):
    """
    ViT+LSTM Network
    Num Params: 3,563,663
    """

    def __init__(self):
        super().__init__()
        self._build_encoder()
        self.decoder = spectral_norm(nn.Linear(4608, 512))
        self.lstm = (nn.LSTM(input_size=517, hidden_size=128, num_layers=3, dropout=0.1))
        self.nn_fc2 = spectral_norm(nn.Linear(128, 3))
        self._build_tail()

    def _pack(self):
        dec = sn_effective_weight(self.decoder)
        pk = {"decoder": dec, "tail": self._pack_tail_tc(dec), "fc2": sn_effective_weight(self.nn_fc2), "lstm": pack_lstm(self.lstm)}
        pk["stage"] = self._pack_stage(pk) if dec.is_cuda else None
        return pk

    # ---- stage-level C ABI (csrc/stages.cu evfly_vit_lstm_forward): the whole forward enqueued by ONE call ----------
    def _pack_stage(self, pk):
        from . import _lib
        w = _lib.VitLstmWeights()
        keep = []

        def p(t):
            t = t.contiguous()
            keep.append(t)
            return t.data_ptr()
        for s, blk in enumerate(self.encoder_blocks):
            bpk = blk.packed()
            c, ln = blk.patchMerge.cn1, blk.patchMerge.layerNorm
            st = w.stage[s]
            st.patch_w, st.patch_b, st.patch_ln_g, st.patch_ln_b = p(bpk["patch_w"]), p(c.bias.float()), p(ln.weight.float()), p(ln.bias.float())
            for l, (attn, lw) in enumerate(zip(blk._attn, bpk["layers"])):
                if lw["attn_fused"] is None or lw["ffn_fused"] is None:
                    return None
                L = st.layer[l]
                L.red_w, L.red_b, L.red_ln_g, L.red_ln_b = p(lw["red_w"]), p(attn.cn1.bias.float()), p(attn.ln1.weight.float()), p(attn.ln1.bias.float())
                L.kv_w, L.kv_b = p(lw["kv"]), p(attn.keyValueExtractor.bias.float())
                L.attn_img, L.attn_bias = p(lw["attn_fused"][0]), p(lw["attn_fused"][1])
                L.ffn_img, L.ffn_bias = p(lw["ffn_fused"][0]), p(lw["ffn_fused"][1])
        w.ds_w, w.ds_b = p(pk["tail"]["ds_w"]), p(pk["tail"]["ds_b"])
        w.dec_w, w.dec_b = p(pk["tail"]["dec"]), p(self.decoder.bias.float())
        for l, (w_ih, b, w_hh_t, pairs, _w_hh) in enumerate(pk["lstm"]):
            if pairs is None or b is None:
                return None
            w.lstm_w_ih[l], w.lstm_b[l], w.lstm_whh_pairs[l], w.lstm_whh_t[l] = p(w_ih), p(b), p(pairs), p(w_hh_t)
        w.fc2_w, w.fc2_b = p(pk["fc2"]), p(self.nn_fc2.bias.float())
        return w, keep

    _stage_ws: dict = {}

    def forward_from_depth(self, depth, desvel, quat, state, n_traj=1, premap_clamp=False):
        """evfly_vit_lstm_forward: depth [N,1,H,W] (any size; premap_clamp: clamp(2 d, 0, 1) first, learner_models.py:634)
        -> (vel [N,3], (h, c)). Batches of N >= 8 frames on the bf16 path."""
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        stage = self.packed()["stage"]
        N, _, H, W = depth.shape
        dev = depth.device
        need = lib.evfly_vit_lstm_workspace_bytes(N)
        key = (dev, torch.cuda.current_stream().cuda_stream)
        ws = LSTMNetVIT._stage_ws.get(key)
        if ws is None or ws.numel() < need:
            ws = LSTMNetVIT._stage_ws[key] = torch.empty((need,), dtype=torch.uint8, device=dev)
        st_shape = (3, 128) if n_traj == 1 else (3, n_traj, 128)
        hT = torch.empty(st_shape, dtype=torch.float32, device=dev)
        cT = torch.empty(st_shape, dtype=torch.float32, device=dev)
        h0 = c0 = None
        if state is not None:
            h0, c0 = to_dev(state[0], dev), to_dev(state[1], dev)
            assert tuple(h0.shape) == st_shape and tuple(c0.shape) == st_shape, "LSTM state shape"
        vel = torch.empty((N, 3), dtype=torch.float32, device=dev)
        dv = to_dev(desvel, dev).reshape(N)
        _lib.check(lib.evfly_vit_lstm_forward(C.byref(stage[0]), _lib.ptr(depth), N, n_traj, H, W, int(premap_clamp), _lib.ptr(dv),
                                              None if quat is None else _lib.ptr(to_dev(quat, dev)), _lib.ptr(h0), _lib.ptr(c0), _lib.ptr(hT), _lib.ptr(cT),
                                              _lib.ptr(vel), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "evfly_vit_lstm_forward")
        return vel, (hT, cT)

    def _stage_usable(self, N):
        return self.precision == 'bf16' and tc.USE_STAGE_ABI and N >= 16 and self.packed()["stage"] is not None

    def forward_trajectories(self, X, n_traj):
        """Extension: n_traj sequences advance together (rows time-major, state [3, n_traj, 128])."""
        return self.forward(X, n_traj=n_traj)

    def forward(self, X, n_traj=1):
        X = _inputs(self, X)
        pk = self.packed()
        N = X[0].shape[0]
        if self._stage_usable(N):
            return self.forward_from_depth(X[0], X[1], X[2], X[3] if len(X) > 3 else None, n_traj)
        seq, feat_out = _meta_concat(512, [(X[1], 1.0, 10.0), (X[2], 1.0, 1.0)], N, X[0].device)   # X[1]/10
        if self.precision == 'bf16' and N >= 16:
            # batches: tail + decoder Linear on the tensor cores, fp32 out into the concat buffer
            feat = self._encode_tc(X[0], pk["tail"])
            tc.gemm_into_f32(feat, pk["tail"]["dec"], self.decoder.bias, seq, 0)
        else:
            feat = self._encode(X[0])
            ops.linear(feat, pk["decoder"], self.decoder.bias, out2d=feat_out)
        state = X[3] if len(X) > 3 else None
        out, h = run_lstm(ops, pk["lstm"], seq, state, 128, smem_weights=self.precision == 'bf16', n_seq=n_traj)
        out = ops.linear(out, pk["fc2"], self.nn_fc2.bias)
        return out, h


class ViT(_ViTEncoder):
    """
    ViT+FC Network
    Num Params: 3,101,199
    """

    def __init__(self):
        super().__init__()
        self._build_encoder()
        self.decoder = nn.Linear(4608, 512)
        self.nn_fc1 = spectral_norm(nn.Linear(517, 256))
        self.nn_fc2 = spectral_norm(nn.Linear(256, 3))
        self._build_tail()

    def _pack(self):
        return {"fc1": sn_effective_weight(self.nn_fc1), "fc2": sn_effective_weight(self.nn_fc2)}

    def forward(self, X):
        X = _inputs(self, X)
        pk = self.packed()
        N = X[0].shape[0]
        feat = self._encode(X[0])
        x, feat_out = _meta_concat(512, [(X[1], 1.0, 10.0), (X[2], 1.0, 1.0)], N, feat.device)
        ops.linear(feat, self.decoder.weight, self.decoder.bias, out2d=feat_out)
        x = ops.linear(x, pk["fc1"], self.nn_fc1.bias, act="leaky_relu")
        return ops.linear(x, pk["fc2"], self.nn_fc2.bias), None


class ConvNet(PackedModule):
    """
    Conv + FC Network
    Num Params: 235,269
    """

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 4, 3, 3)
        self.conv2 = nn.Conv2d(4, 10, 3, 2)
        self.avgpool = nn.AvgPool2d(kernel_size=3, stride=1)
        self.maxpool = nn.MaxPool2d(2, 1)
        self.bn1 = nn.BatchNorm2d(4)
        self.fc0 = nn.Linear(845, 256, bias=False)
        self.fc1 = nn.Linear(256, 64, bias=False)
        self.fc2 = nn.Linear(64, 32, bias=False)
        self.fc3 = nn.Linear(32, 3)

    def _pack(self):
        return {"bn1": bn_affine(self.bn1)}

    def forward(self, X):
        X = _inputs(self, X)
        pk = self.packed()
        N = X[0].shape[0]
        # -maxpool(-bn1(relu(conv1(x))))   (vitfly_models.py:58)
        x = ops.conv2d(X[0], self.conv1.weight, self.conv1.bias, stride=3, act="relu", post=pk["bn1"])
        x = ops.pool2d(x, 2, 1, "max", negate_in=True, negate_out=True)
        x = ops.conv2d(x, self.conv2.weight, self.conv2.bias, stride=2, act="relu")
        x = ops.pool2d(x, 3, 1, "avg").view(N, -1)
        feat, conv_out = _meta_concat(x.shape[1], [(X[1], 0.1, 1.0), (X[2], 1.0, 1.0)], N, x.device)   # X[1]*0.1
        ops.map4d(x, conv_out)
        x = ops.linear(feat, self.fc0.weight, None, act="leaky_relu")
        x = ops.linear(x, self.fc1.weight, None, act="leaky_relu")
        x = ops.linear(x, self.fc2.weight, None, act="tanh")
        return ops.linear(x, self.fc3.weight, self.fc3.bias), None


class LSTMNet(PackedModule):
    """
    LSTM + FC Network
    Num Params: 2,949,937
    """

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 4, 5, stride=3, padding=1)
        self.conv2 = nn.Conv2d(4, 10, 3, stride=2, padding=0)
        self.avgpool = nn.AvgPool2d(kernel_size=3, stride=1)
        self.maxpool = nn.MaxPool2d(3, 1)
        self.bn1 = nn.BatchNorm2d(4)
        self.bn2 = nn.BatchNorm2d(10)
        self.lstm = LSTM(input_size=665, hidden_size=395, num_layers=2, dropout=0.15, bias=False)
        self.fc1 = spectral_norm(nn.Linear(395, 64))
        self.fc2 = spectral_norm(nn.Linear(64, 16))
        self.fc3 = spectral_norm(nn.Linear(16, 3))

    def _pack(self):
        return {"bn1": bn_affine(self.bn1), "bn2": bn_affine(self.bn2), "lstm": pack_lstm(self.lstm),
                "fc": [sn_effective_weight(m) for m in (self.fc1, self.fc2, self.fc3)]}

    def forward(self, X):
        X = _inputs(self, X)
        pk = self.packed()
        N = X[0].shape[0]
        x = ops.conv2d(X[0], self.conv1.weight, self.conv1.bias, stride=3, pad=1, act="relu", post=pk["bn1"])
        x = ops.pool2d(x, 3, 1, "max", negate_in=True, negate_out=True)
        x = ops.conv2d(x, self.conv2.weight, self.conv2.bias, stride=2, act="relu", post=pk["bn2"])
        x = ops.pool2d(x, 3, 1, "avg").view(N, -1)
        seq, conv_out = _meta_concat(x.shape[1], [(X[1], 0.1, 1.0), (X[2], 1.0, 1.0)], N, x.device)
        ops.map4d(x, conv_out)
        out, h = run_lstm(ops, pk["lstm"], seq, X[3] if len(X) > 3 else None, 395)
        out = ops.linear(out, pk["fc"][0], self.fc1.bias, act="leaky_relu")
        out = ops.linear(out, pk["fc"][1], self.fc2.bias, act="leaky_relu")
        return ops.linear(out, pk["fc"][2], self.fc3.bias), h


class UNetConvLSTMNet(PackedModule):
    """
    UNet+LSTM Network
    Num Params: 2,955,822
    """

    def __init__(self):
        super().__init__()
        self.unet_e11 = nn.Conv2d(1, 4, kernel_size=3, padding=1)
        self.unet_e12 = nn.Conv2d(4, 4, kernel_size=3, padding=1)
        self.unet_pool1 = nn.MaxPool2d(kernel_size=2, stride=3,)
        self.unet_e21 = nn.Conv2d(4, 8, kernel_size=3, padding=1)
        self.unet_e22 = nn.Conv2d(8, 8, kernel_size=3, padding=1)
        self.unet_pool2 = nn.MaxPool2d(kernel_size=2, stride=2,)
        self.unet_e31 = nn.Conv2d(8, 16, kernel_size=3, padding=1)
        self.unet_e32 = nn.Conv2d(16, 16, kernel_size=3, padding=1)
        self.unet_upconv1 = nn.ConvTranspose2d(16, 8, kernel_size=2, stride=2,)
        self.unet_d11 = nn.Conv2d(16, 8, kernel_size=3, padding=1)
        self.unet_d12 = nn.Conv2d(8, 8, kernel_size=3, padding=1)
        self.unet_upconv2 = nn.ConvTranspose2d(8, 4, kernel_size=3, stride=3,)
        self.unet_d21 = nn.Conv2d(8, 4, kernel_size=3, padding=1)
        self.unet_d22 = nn.Conv2d(4, 4, kernel_size=3, padding=1)
        self.unet_out = nn.Conv2d(4, 1, kernel_size=1)
        self.conv_conv1 = nn.Conv2d(2, 4, 5, 3)
        self.conv_conv2 = nn.Conv2d(4, 10, 5, 2)
        self.conv_avgpool = nn.AvgPool2d(kernel_size=2, stride=1)
        self.conv_maxpool = nn.MaxPool2d(2, 1)
        self.conv_bn1 = nn.BatchNorm2d(4)
        self.lstm = LSTM(input_size=3065, hidden_size=200, num_layers=2, dropout=0.15, bias=False)
        self.nn_fc1 = torch.nn.utils.spectral_norm(nn.Linear(200, 64))
        self.nn_fc2 = torch.nn.utils.spectral_norm(nn.Linear(64, 32))
        self.nn_fc3 = torch.nn.utils.spectral_norm(nn.Linear(32, 3))

    def _pack(self):
        # relu(bn(conv)) folds the BatchNorm into the conv: w*scale, bias*scale+shift
        scale, shift = bn_affine(self.conv_bn1)
        w = (self.conv_conv1.weight * scale.view(-1, 1, 1, 1)).contiguous()
        b = (self.conv_conv1.bias * scale + shift).contiguous()
        return {"conv1": (w, b), "lstm": pack_lstm(self.lstm),
                "up1": pack_conv_transpose(self.unet_upconv1), "up2": pack_conv_transpose(self.unet_upconv2),
                "fc": [sn_effective_weight(m) for m in (self.nn_fc1, self.nn_fc2, self.nn_fc3)]}

    def forward(self, X):
        X = _inputs(self, X)
        pk = self.packed()
        img = X[0]
        N = img.shape[0]
        dev = img.device
        c = lambda m, x, **kw: ops.conv2d(x, m.weight, m.bias, **kw)
        e1 = c(self.unet_e12, c(self.unet_e11, img, pad=1, act="relu"), pad=1, act="relu")          # [N,4,60,90]
        e2 = c(self.unet_e22, c(self.unet_e21, ops.pool2d(e1, 2, 3), pad=1, act="relu"), pad=1, act="relu")   # [N,8,20,30]
        e3 = c(self.unet_e32, c(self.unet_e31, ops.pool2d(e2, 2, 2), pad=1, act="relu"), pad=1, act="relu")   # [N,16,10,15]
        cat1 = torch.empty((N, 16, 20, 30), dtype=torch.float32, device=dev)
        conv_transpose(e3, pk["up1"], self.unet_upconv1.bias, cat1[:, :8])
        ops.map4d(e2, cat1[:, 8:])
        d1 = c(self.unet_d12, c(self.unet_d11, cat1, pad=1, act="relu"), pad=1, act="relu")
        cat2 = torch.empty((N, 8, 60, 90), dtype=torch.float32, device=dev)
        conv_transpose(d1, pk["up2"], self.unet_upconv2.bias, cat2[:, :4])
        ops.map4d(e1, cat2[:, 4:])
        d2 = c(self.unet_d22, c(self.unet_d21, cat2, pad=1, act="relu"), pad=1, act="relu")
        x_conv = torch.empty((N, 2, 60, 90), dtype=torch.float32, device=dev)
        ops.map4d(img, x_conv[:, :1])
        c(self.unet_out, d2, out_view=x_conv[:, 1:])
        y = ops.conv2d(x_conv, pk["conv1"][0], pk["conv1"][1], stride=3, act="relu")
        y = ops.pool2d(y, 2, 1, "max", negate_in=True, negate_out=True)
        y = ops.pool2d(c(self.conv_conv2, y, stride=2, act="relu"), 2, 1, "avg").view(N, -1)
        e3f = e3.view(N, -1)
        seq, head = _meta_concat(y.shape[1] + e3f.shape[1], [(X[1], 0.1, 1.0), (X[2], 1.0, 1.0)], N, dev)
        ops.map4d(y, head[:, :y.shape[1]])
        ops.map4d(e3f, head[:, y.shape[1]:])
        out, h = run_lstm(ops, pk["lstm"], seq, X[3] if len(X) > 3 else None, 200)
        out = ops.linear(out, pk["fc"][0], self.nn_fc1.bias, act="leaky_relu")
        out = ops.linear(out, pk["fc"][1], self.nn_fc2.bias, act="leaky_relu")
        return ops.linear(out, pk["fc"][2], self.nn_fc3.bias), h


# ---- ConvTranspose2d with kernel_size == stride (every use in the reference) -------------------
def pack_conv_transpose(m: nn.ConvTranspose2d):
    """[Cin,Cout,k,k] -> [k,k,Cout,Cin,1,1]: one 1x1-conv weight matrix per output phase."""
    k = m.kernel_size[0]
    assert m.kernel_size == m.stride and m.padding == (0, 0) and m.kernel_size[0] == m.kernel_size[1]
    return m.weight.permute(2, 3, 1, 0).contiguous().view(k, k, m.out_channels, m.in_channels, 1, 1)


def conv_transpose(x, packed_w, bias, out_view):
    """Each output pixel (k*ih+a, k*iw+b) depends on input pixel (ih,iw) only: k*k 1x1 convs whose
    outputs interleave through strides into out_view [N,Cout,k*H,k*W] (e.g. a concat slice)."""
    k = packed_w.shape[0]
    for a in range(k):
        for b in range(k):
            ops.conv2d(x, packed_w[a, b], bias, out_view=out_view[:, :, a::k, b::k])
    return out_view
