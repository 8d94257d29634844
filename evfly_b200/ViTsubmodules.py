"""Drop-in for evfly's learner/ViTsubmodules.py (same class names, constructor signatures,
parameter names and forward signatures), computed by libevfly_b200 kernels.

Token tensors are [B, N, C] contiguous; every flatten/transpose/permute of the reference
(ViTsubmodules.py:31-32,65-72,108-114,147) is a strided VIEW handed to the conv kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops, tc
from ._modbase import PackedModule


def _bchw_view(tok: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """[B, H*W, C] tokens seen as [B, C, H, W] (no copy)."""
    B, N, C = tok.shape
    return tok.view(B, H, W, C).permute(0, 3, 1, 2)


class OverlapPatchMerging(PackedModule):
    def __init__(self, in_channels, out_channels, patch_size, stride, padding):
        super().__init__()
        self.cn1 = nn.Conv2d(in_channels, out_channels, kernel_size=patch_size, stride=stride, padding=padding)
        self.layerNorm = nn.LayerNorm(out_channels)

    def forward(self, patches):
        """(B,C,H,W) -> tokens (B, H'*W', C_out), H', W'   (ViTsubmodules.py:21-34)"""
        B, _, H, W = patches.shape
        c = self.cn1
        k, s, p = c.kernel_size[0], c.stride[0], c.padding[0]
        OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        tok = torch.empty((B, OH * OW, c.out_channels), dtype=torch.float32, device=patches.device)
        ops.conv2d(patches, c.weight, c.bias, stride=s, pad=p, out_view=_bchw_view(tok, OH, OW))
        ln = self.layerNorm
        return ops.layernorm(tok, ln.weight, ln.bias, ln.eps), OH, OW


class EfficientSelfAttention(PackedModule):
    def __init__(self, channels, reduction_ratio, num_heads):
        super().__init__()
        assert channels % num_heads == 0, f"channels {channels} should be divided by num_heads {num_heads}."
        self.heads = num_heads
        self.cn1 = nn.Conv2d(in_channels=channels, out_channels=channels, kernel_size=reduction_ratio, stride=reduction_ratio)
        self.ln1 = nn.LayerNorm(channels)
        self.keyValueExtractor = nn.Linear(channels, channels * 2)
        self.query = nn.Linear(channels, channels)
        self.smax = nn.Softmax(dim=-1)
        self.finalLayer = nn.Linear(channels, channels)

    def forward(self, x, H, W, residual=None):
        """(B,N,C) -> (B,N,C); with `residual` the sum residual + attn(x) is returned (the
        caller's `x = x + attn(x)`, ViTsubmodules.py:144, fused into the last GEMM's epilogue)."""
        B, N, C = x.shape
        r = self.cn1.stride[0]
        h2, w2 = (H - r) // r + 1, (W - r) // r + 1
        red = torch.empty((B, h2 * w2, C), dtype=torch.float32, device=x.device)
        ops.conv2d(_bchw_view(x, H, W), self.cn1.weight, self.cn1.bias, stride=r, out_view=_bchw_view(red, h2, w2))
        red = ops.layernorm(red, self.ln1.weight, self.ln1.bias, self.ln1.eps)
        kv = ops.linear(red.view(-1, C), self.keyValueExtractor.weight, self.keyValueExtractor.bias).view(B, h2 * w2, 2 * C)
        q = ops.linear(x.view(-1, C), self.query.weight, self.query.bias).view(B, N, C)
        att = ops.attention_small(q, kv, self.heads)
        out = ops.linear(att.view(-1, C), self.finalLayer.weight, self.finalLayer.bias,
                         res2d=None if residual is None else residual.view(-1, C))
        return out.view(B, N, C)


class MixFFN(PackedModule):
    def __init__(self, channels, expansion_factor):
        super().__init__()
        expanded_channels = channels * expansion_factor
        self.mlp1 = nn.Linear(channels, expanded_channels)
        self.depthwise = nn.Conv2d(expanded_channels, expanded_channels, kernel_size=3, padding='same', groups=channels)
        self.gelu = nn.GELU()
        self.mlp2 = nn.Linear(expanded_channels, channels)

    def forward(self, x, H, W, residual=None):
        """(B,N,C) -> (B,N,C): Linear, grouped 3x3 conv ('same'), exact GELU, Linear
        (ViTsubmodules.py:105-120). GELU rides in the conv epilogue, the residual in mlp2's."""
        B, N, C = x.shape
        Ce = self.mlp1.out_features
        y1 = ops.linear(x.view(-1, C), self.mlp1.weight, self.mlp1.bias).view(B, N, Ce)
        y2 = torch.empty_like(y1)
        ops.conv2d(_bchw_view(y1, H, W), self.depthwise.weight, self.depthwise.bias, pad=1, groups=self.depthwise.groups,
                   act="gelu", out_view=_bchw_view(y2, H, W))
        out = ops.linear(y2.view(-1, Ce), self.mlp2.weight, self.mlp2.bias,
                         res2d=None if residual is None else residual.view(-1, C))
        return out.view(B, N, C)


class MixTransformerEncoderLayer(PackedModule):
    def __init__(self, in_channels, out_channels, patch_size, stride, padding,
                 n_layers, reduction_ratio, num_heads, expansion_factor):
        super().__init__()
        self.patchMerge = OverlapPatchMerging(in_channels, out_channels, patch_size, stride, padding)
        self._attn = nn.ModuleList([EfficientSelfAttention(out_channels, reduction_ratio, num_heads) for _ in range(n_layers)])
        self._ffn = nn.ModuleList([MixFFN(out_channels, expansion_factor) for _ in range(n_layers)])
        self._lNorm = nn.ModuleList([nn.LayerNorm(out_channels) for _ in range(n_layers)])

    precision = 'fp32'     # 'bf16': Linear layers on the tensor cores, bf16 NHWC tokens (evfly_b200.set_precision)

    def _pack(self):
        c = self.patchMerge.cn1
        pk = {"patch_w": tc.pack_conv_kc(c.weight), "layers": []}
        for attn, ffn, ln in zip(self._attn, self._ffn, self._lNorm):
            fused = None
            if ffn.mlp1.out_features == 8 * ffn.mlp1.in_features and ffn.mlp1.in_features in (32, 64) and ffn.depthwise.weight.is_cuda:
                fused = tc.pack_vit_ffn(ffn.mlp1.weight, ffn.mlp1.bias, ffn.depthwise.weight, ffn.depthwise.bias, ffn.mlp2.weight, ffn.mlp2.bias,
                                        ln.weight, ln.bias)
            attn_fused = None
            if attn.query.in_features in (32, 64) and attn.query.in_features // attn.heads == 32 and attn.query.weight.is_cuda:
                attn_fused = tc.pack_vit_attn(attn.query.weight, attn.query.bias, attn.finalLayer.weight, attn.finalLayer.bias)
            pk["layers"].append({
                "ffn_fused": fused, "attn_fused": attn_fused,
                "red_w": tc.pack_conv_kc(attn.cn1.weight),
                "kv": tc.pack_conv1x1_weight(attn.keyValueExtractor.weight), "q": tc.pack_conv1x1_weight(attn.query.weight),
                "final": tc.pack_conv1x1_weight(attn.finalLayer.weight),
                "mlp1": tc.pack_conv1x1_weight(ffn.mlp1.weight), "mlp2": tc.pack_conv1x1_weight(ffn.mlp2.weight)})
        return pk

    def encode_bf16(self, x, x_is_f32_nchw, B, H, W):
        """bf16 path of forward(): x is the fp32 NCHW depth image (stage 1) or the previous stage's
        bf16 tokens seen as NHWC [B,H,W,Cin]. Returns (tokens bf16 [B,N,C], H', W')."""
        pk = self.packed()
        pm, c = self.patchMerge, self.patchMerge.cn1
        C = c.out_channels
        tok, H2, W2 = tc.patch_embed_ln(x, x_is_f32_nchw, pk["patch_w"], c.bias, pm.layerNorm.weight, pm.layerNorm.bias, B, H, W,
                                        c.in_channels, C, c.kernel_size[0], c.stride[0], c.padding[0], pm.layerNorm.eps)
        N = H2 * W2
        for attn, ffn, ln, w in zip(self._attn, self._ffn, self._lNorm, pk["layers"]):
            r = attn.cn1.stride[0]
            # spatial-reduction attention: k=s=r conv + LayerNorm fused, K/V and Q projections, few-key softmax
            red, h2, w2 = tc.patch_embed_ln(tok, False, w["red_w"], attn.cn1.bias, attn.ln1.weight, attn.ln1.bias, B, H2, W2, C, C, r, r, 0, attn.ln1.eps)
            kv = tc.gemm_tokens(red.view(-1, C), w["kv"], attn.keyValueExtractor.bias).view(B, h2 * w2, 2 * C)
            if w["attn_fused"] is not None and tc.FUSED_ATTN and h2 * w2 <= 8:
                tok = tc.vit_attn(tok, kv, w["attn_fused"][0], w["attn_fused"][1], attn.heads)                       # x + attn(x), one launch
            else:
                q = tc.gemm_tokens(tok.view(-1, C), w["q"], attn.query.bias).view(B, N, C)
                att = tc.attention_small_bf16(q, kv, attn.heads)
                tok = tc.gemm_tokens(att.view(-1, C), w["final"], attn.finalLayer.bias, res_bf16=tok).view(B, N, C)  # x + attn(x)
            # MixFFN + residual + LayerNorm: one launch, the 8C-wide activation never leaves the SM (csrc/vit_fused.cu)
            if w["ffn_fused"] is not None and B >= tc.FUSED_FFN_MIN_BATCH and (H2, W2, C) in tc.FUSED_FFN_SHAPES:
                tok = tc.vit_ffn(tok, w["ffn_fused"][0], w["ffn_fused"][1], B, H2, W2, ln.eps)
                continue
            # per-op path (small batches): Linear -> grouped 3x3 + GELU -> Linear, residual in the last epilogue
            y1 = tc.gemm_tokens(tok.view(-1, C), w["mlp1"], ffn.mlp1.bias)
            y2 = tc.dwconv3x3_gelu(y1.view(B, H2, W2, -1), ffn.depthwise.weight, ffn.depthwise.bias)
            tok = tc.gemm_tokens(y2.view(B * N, -1), w["mlp2"], ffn.mlp2.bias, res_bf16=tok).view(B, N, C)          # x + ffn(x)
            tok = tc.layernorm_bf16(tok, ln.weight, ln.bias, ln.eps)
        return tok, H2, W2

    def forward(self, x):
        """(B,C,H,W) -> (B,C',H',W') (a channel-last view; ViTsubmodules.py:132-148)."""
        self._check_inference()
        B = x.shape[0]
        if self.precision == 'bf16' and x.shape[1] == 1:
            tok, H, W = self.encode_bf16(x.contiguous(), True, B, x.shape[2], x.shape[3])
            return tc.grid_to_nchw(tok.view(B, H, W, -1), H, W)
        tok, H, W = self.patchMerge(x)
        for i in range(len(self._attn)):
            tok = self._attn[i].forward(tok, H, W, residual=tok)
            tok = self._ffn[i].forward(tok, H, W, residual=tok)
            ln = self._lNorm[i]
            tok = ops.layernorm(tok, ln.weight, ln.bias, ln.eps)
        return _bchw_view(tok, H, W)
