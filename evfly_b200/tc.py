"""Typed wrappers of the bf16 tensor-core entry points (include/evfly_b200.h, bf16 section).

A `Grid` is a bf16 NHWC activation on a pitch-preserving grid: data [N, Hp, Wp, C] plus the valid
extent (vh, vw). See evfly_b200/csrc/tc_conv_bf16.cu for why this layout turns a 3x3 valid conv
into nine shifted GEMMs.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib

BF16 = torch.bfloat16
USE_HALO = True      # small-channel 3x3 convs through the halo-reuse kernel
FUSE_POOL = True     # MaxPool2d(2) in the epilogue of the halo-reuse kernel
PERSISTENT_SCAN = True   # ConvLSTM recurrence as one persistent launch with a grid barrier per step
COMPACT_GRIDS = True     # wide (generic-kernel) 3x3 convs write compact grids: no don't-care rows in the next conv / the ConvLSTM
SKIP_ROWS = True         # fused-pool convs whose output only feeds the 'interp' skip write just the rows that resize samples
FUSED_SCAN = True        # ... with the x half of the gate conv inside the step (no fp32 x-gate tensor); needs PERSISTENT_SCAN


@dataclass
class Grid:
    data: torch.Tensor   # bf16 [N, Hp, Wp, C] contiguous
    vh: int
    vw: int

    @property
    def N(self): return self.data.shape[0]
    @property
    def Hp(self): return self.data.shape[1]
    @property
    def Wp(self): return self.data.shape[2]
    @property
    def C(self): return self.data.shape[3]
    @property
    def rows(self): return self.data.shape[0] * self.data.shape[1] * self.data.shape[2]


def new_grid(N, Hp, Wp, Cc, vh, vw, device) -> Grid:
    return Grid(torch.empty((N, Hp, Wp, Cc), dtype=BF16, device=device), vh, vw)


def pack_conv3x3_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] fp32 -> bf16 [Cout, 9*Cin] with K index = tap*Cin + ci (tap = kh*3+kw)."""
    Cout, Cin = w.shape[:2]
    return w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).to(BF16).contiguous()


def pack_conv1x1_weight(w: torch.Tensor) -> torch.Tensor:
    return w.reshape(w.shape[0], -1).to(BF16).contiguous()


def pack_convt2x2_weight(w: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d weight [Cin, Cout, 2, 2] -> bf16 [4*Cout, Cin], row = (2a+b)*Cout + co."""
    Cin, Cout = w.shape[:2]
    return w.permute(2, 3, 1, 0).reshape(4 * Cout, Cin).to(BF16).contiguous()


def _call_scan(*args):
    _lib.check(_lib.load().evfly_convlstm_scan_bf16(*args, _lib.stream_ptr()), "evfly_convlstm_scan_bf16")


def _call_scan_fused(*args) -> bool:
    """False when the device cannot keep the persistent grid co-resident (the caller then takes the two-kernel path)."""
    rc = _lib.load().evfly_convlstm_scan_fused_bf16(*args, _lib.stream_ptr())
    if rc == _lib.ERR_UNSUPPORTED:
        return False
    _lib.check(rc, "evfly_convlstm_scan_fused_bf16")
    return True


def _call_halo(*args, compact=False):
    if compact:
        _lib.check(_lib.load().evfly_tc_conv3x3_halo_compact_bf16(*args, _lib.stream_ptr()), "evfly_tc_conv3x3_halo_compact_bf16")
    else:
        _lib.check(_lib.load().evfly_tc_conv3x3_halo_bf16(*args, _lib.stream_ptr()), "evfly_tc_conv3x3_halo_bf16")


def _call_halo_pool(*args):
    """15 arguments: evfly_tc_conv3x3_halo_pool_bf16; a 16th (skip_OH): the _rows variant."""
    if len(args) == 16:
        _lib.check(_lib.load().evfly_tc_conv3x3_halo_pool_rows_bf16(*args, _lib.stream_ptr()), "evfly_tc_conv3x3_halo_pool_rows_bf16")
    else:
        _lib.check(_lib.load().evfly_tc_conv3x3_halo_pool_bf16(*args, _lib.stream_ptr()), "evfly_tc_conv3x3_halo_pool_bf16")


def _halo_ok(cin, cout):
    return USE_HALO and ((cin in (32, 64) and cout in (32, 64)) or (cin == 64 and cout == 128) or (cin == 128 and cout in (64, 128, 256)))


def conv3x3_pool(g: Grid, w_packed, bias, relu=True, skip_rows=None):
    """conv3x3 followed by MaxPool2d(2): returns (conv output, pooled output). The pool is fused into the conv
    epilogue on the halo path, a separate pass otherwise; the two are bit-identical. skip_rows = OH (halo path only): the conv
    output will only be read by resize_bilinear_into(..., OH, ...); rows that resize does not sample are not written."""
    Cout = w_packed.shape[0]
    if not (_halo_ok(g.C, Cout) and FUSE_POOL):
        y = conv3x3(g, w_packed, bias, relu=relu, compact=True)
        return y, maxpool2x2(y)
    out = new_grid(g.N, g.Hp, g.Wp, Cout, g.vh - 2, g.vw - 2, g.data.device)
    ph, pw = (g.vh - 2) // 2, (g.vw - 2) // 2
    pooled = new_grid(g.N, ph, pw, Cout, ph, pw, g.data.device)
    args = (g.data.data_ptr(), w_packed.data_ptr(), _lib.ptr(bias), out.data.data_ptr(), pooled.data.data_ptr(), g.N, g.Hp, g.Wp,
            g.vh, g.vw, g.C, Cout, int(relu), pooled.Hp, pooled.Wp)
    _call_halo_pool(*args, int(skip_rows)) if (skip_rows and SKIP_ROWS) else _call_halo_pool(*args)
    return out, pooled


def conv3x3_same(x_nhwc: torch.Tensor, w_packed, bias, relu=False) -> torch.Tensor:
    """Conv2d(k=3, padding=1) on a dense bf16 NHWC tensor [N,H,W,Cin] (Cin in {32,64}) -> [N,H,W,Cout]."""
    N, H, W, Cin = x_nhwc.shape
    Cout = w_packed.shape[0]
    assert x_nhwc.is_contiguous() and x_nhwc.dtype == BF16
    out = torch.empty((N, H, W, Cout), dtype=BF16, device=x_nhwc.device)
    _lib.check(_lib.load().evfly_tc_conv3x3_same_bf16(x_nhwc.data_ptr(), w_packed.data_ptr(), _lib.ptr(bias), out.data_ptr(), N, H, W, Cin, Cout,
                                                       int(relu), _lib.stream_ptr()), "evfly_tc_conv3x3_same_bf16")
    return out


def shuffle_upsample_cat(t2, H2, W2, t1, H1, W1, ld) -> torch.Tensor:
    """cat([PixelShuffle(2)(t2), Upsample((2*H2,2*W2), align_corners=True)(t1)]) as bf16 NHWC [B,2*H2,2*W2,ld] (zero-padded channels)."""
    B = t2.shape[0]
    C2, C1 = t2.shape[-1], t1.shape[-1]
    out = torch.empty((B, 2 * H2, 2 * W2, ld), dtype=BF16, device=t2.device)
    _lib.check(_lib.load().evfly_shuffle_upsample_cat_bf16(t2.data_ptr(), H2, W2, C2, t1.data_ptr(), H1, W1, C1, out.data_ptr(), B, ld,
                                                            _lib.stream_ptr()), "evfly_shuffle_upsample_cat_bf16")
    return out


def _call(a: _lib.TcConvArgs):
    _lib.check(_lib.load().evfly_tc_conv_bf16(C.byref(a), _lib.stream_ptr()), "evfly_tc_conv_bf16")


def conv3x3(g: Grid, w_packed, bias, relu=True, out: Grid | None = None, compact=False) -> Grid:
    """3x3 valid conv (+bias, ReLU) on the grid; the result keeps the pitch, valid extent - 2. compact=True: the result is
    written on a grid whose pitch IS its valid extent, so a consumer that computes every row of the grid it is given (the
    generic kernel, the transposed convs, the ConvLSTM) has no don't-care rows to compute."""
    Cout = w_packed.shape[0]
    compact = bool(compact) and COMPACT_GRIDS and out is None
    if out is None:
        out = (new_grid(g.N, g.vh - 2, g.vw - 2, Cout, g.vh - 2, g.vw - 2, g.data.device) if compact else
               new_grid(g.N, g.Hp, g.Wp, Cout, g.vh - 2, g.vw - 2, g.data.device))
    if _halo_ok(g.C, Cout):
        # small-channel layers: halo reuse from shared memory, weights resident (tc_conv_halo.cu)
        args = (g.data.data_ptr(), w_packed.data_ptr(), _lib.ptr(bias), out.data.data_ptr(), g.N, g.Hp, g.Wp, g.vh, g.vw, g.C, Cout, int(relu))
        _call_halo(*args, compact=True) if compact else _call_halo(*args)
        return out
    a = _lib.TcConvArgs()
    a.x, a.w, a.bias, a.out = g.data.data_ptr(), w_packed.data_ptr(), _lib.ptr(bias), out.data.data_ptr()
    a.M_rows, a.out_ld = g.rows, Cout
    a.Cin, a.n_rows, a.taps, a.w_pitch, a.relu, a.out_c0 = g.C, Cout, 9, g.Wp, int(relu), 0
    if compact:
        a.flags, a.Hp, a.Wp, a.valid_h, a.valid_w = _lib.TC_COMPACT, g.Hp, g.Wp, g.vh - 2, g.vw - 2
    _call(a)
    return out


def gemm(x2d, w_packed, bias=None, relu=False, out2d=None, out_f32=None, res_f32=None):
    """[M, K] bf16 @ [N, K]^T -> bf16 [M, N] (or fp32 when out_f32 is given); res_f32 [M, N] added."""
    M, K = x2d.shape
    Nn = w_packed.shape[0]
    assert x2d.is_contiguous() and x2d.dtype == BF16 and w_packed.shape[1] == K
    a = _lib.TcConvArgs()
    a.x, a.w, a.bias = x2d.data_ptr(), w_packed.data_ptr(), _lib.ptr(bias)
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.is_contiguous()
        a.out_f32 = out_f32.data_ptr()
        ret = out_f32
    else:
        if out2d is None:
            out2d = torch.empty((M, Nn), dtype=BF16, device=x2d.device)
        a.out = out2d.data_ptr()
        ret = out2d
    a.res_f32 = None if res_f32 is None else res_f32.data_ptr()
    a.M_rows, a.out_ld = M, Nn
    a.Cin, a.n_rows, a.taps, a.w_pitch, a.relu, a.out_c0 = K, Nn, 1, 0, int(relu), 0
    _call(a)
    return ret


def pack_convlstm_gate_weight(w2d: torch.Tensor) -> torch.Tensor:
    """[4*Ch, K] with rows gate-major (i,f,o,g blocks, convlstm.py:44) -> bf16 rows interleaved n = 4*ch + gate,
    so that one 32-column epilogue slab holds all four gates of 8 channels."""
    G, K = w2d.shape
    Ch = G // 4
    return w2d.reshape(4, Ch, K).permute(1, 0, 2).reshape(G, K).to(BF16).contiguous()


def convlstm_step(h_prev, wh_packed, gx_t, c_f32, h_out):
    """One ConvLSTM step in ONE launch: gates = h_prev @ Wh^T + gx_t (fp32, gate-interleaved), cell update in
    the epilogue. h_prev / h_out bf16 [P,Ch], c fp32 [P,Ch] in place."""
    P, Ch = c_f32.shape
    a = _lib.TcConvArgs()
    a.x, a.w, a.res_f32 = h_prev.data_ptr(), wh_packed.data_ptr(), gx_t.data_ptr()
    a.lstm_c, a.lstm_h = c_f32.data_ptr(), h_out.data_ptr()
    a.M_rows, a.out_ld = P, 4 * Ch
    a.Cin, a.n_rows, a.taps, a.w_pitch, a.relu, a.out_c0 = Ch, 4 * Ch, 1, 0, 0, 0
    _call(a)


def convlstm_scan(h_all, wh_packed, gx, c_f32, T, P, Ch):
    """h_all bf16 [(T+1), P, Ch] (block 0 = h_0), gx fp32 [T*P, 4Ch], c fp32 [P,Ch]: the whole recurrence in one call."""
    sync = torch.empty((1,), dtype=torch.int64, device=c_f32.device) if PERSISTENT_SCAN else None
    _call_scan(h_all.data_ptr(), wh_packed.data_ptr(), _lib.ptr(gx), _lib.ptr(c_f32), T, P, Ch, _lib.ptr(sync))


def convlstm_scan_fused(x_rows, wx_packed, h_all, wh_packed, c_f32, T, P, Ch) -> bool:
    """x_rows bf16 [T*P, Cx] (time-major), h_all bf16 [(T+1), P, Ch] (block 0 = h_0), c fp32 [P,Ch]: the whole recurrence,
    x-gates included, in one persistent launch. Returns False if that launch is not possible here (nothing was run)."""
    if not (PERSISTENT_SCAN and FUSED_SCAN) or T < 1 or x_rows.shape[1] % 64 or Ch % 64:
        return False
    sync = torch.empty((1,), dtype=torch.int64, device=c_f32.device)
    return _call_scan_fused(x_rows.data_ptr(), wx_packed.data_ptr(), h_all.data_ptr(), wh_packed.data_ptr(), _lib.ptr(c_f32), T, P, x_rows.shape[1], Ch,
                            _lib.ptr(sync))


def conv_transpose2x2(g: Grid, w_packed, bias, out_data: torch.Tensor, out_c0: int):
    """ConvTranspose2d(k=2,s=2) of the valid region into channels [out_c0, out_c0+Cout) of the
    compact grid out_data [N, 2*vh, 2*vw, Ctot]."""
    Cout = w_packed.shape[0] // 4
    assert tuple(out_data.shape[:3]) == (g.N, 2 * g.vh, 2 * g.vw)
    a = _lib.TcConvArgs()
    a.x, a.w, a.bias, a.out = g.data.data_ptr(), w_packed.data_ptr(), _lib.ptr(bias), out_data.data_ptr()
    a.M_rows, a.out_ld = g.rows, out_data.shape[3]
    a.Cin, a.n_rows, a.taps, a.w_pitch, a.relu, a.out_c0 = g.C, 4 * Cout, 1, 0, 0, out_c0
    a.convt, a.Hp, a.Wp, a.valid_h, a.valid_w, a.cout_t = 1, g.Hp, g.Wp, g.vh, g.vw, Cout
    _call(a)


def stem_conv3x3(x_nchw_f32, w, bias, fma=False) -> Grid:
    """UNet stem on the tensor cores (image and weights rounded to bf16); fma=True: fp32 CUDA-core version."""
    N, Cin, H, W = x_nchw_f32.shape
    out = new_grid(N, H, W, 32, H - 2, W - 2, x_nchw_f32.device)
    lib = _lib.load()
    fn = lib.evfly_stem_conv3x3_fma_bf16 if fma else lib.evfly_stem_conv3x3_bf16
    _lib.check(fn(_lib.ptr(x_nchw_f32), _lib.ptr(w), _lib.ptr(bias), out.data.data_ptr(), N, Cin, H, W, _lib.stream_ptr()),
               "evfly_stem_conv3x3_bf16")
    return out


def maxpool2x2(g: Grid) -> Grid:
    out = new_grid(g.N, g.vh // 2, g.vw // 2, g.C, g.vh // 2, g.vw // 2, g.data.device)
    _lib.check(_lib.load().evfly_maxpool2x2_nhwc_bf16(g.data.data_ptr(), out.data.data_ptr(), g.N, g.Hp, g.Wp, g.vh, g.vw, g.C,
                                                       _lib.stream_ptr()), "evfly_maxpool2x2_nhwc_bf16")
    return out


def resize_bilinear_into(g: Grid, OH, OW, out_data, out_c0):
    assert tuple(out_data.shape[:3]) == (g.N, OH, OW)
    _lib.check(_lib.load().evfly_resize_bilinear_nhwc_bf16(g.data.data_ptr(), out_data.data_ptr(), g.N, g.Hp, g.Wp, g.vh, g.vw, g.C,
                                                            OH, OW, out_data.shape[3], out_c0, _lib.stream_ptr()),
               "evfly_resize_bilinear_nhwc_bf16")


def crop_into(g: Grid, h0, w0, OH, OW, out_data, out_c0):
    assert tuple(out_data.shape[:3]) == (g.N, OH, OW)
    _lib.check(_lib.load().evfly_crop_nhwc_bf16(g.data.data_ptr(), out_data.data_ptr(), g.N, g.Hp, g.Wp, g.C, h0, w0, OH, OW,
                                                 out_data.shape[3], out_c0, _lib.stream_ptr()), "evfly_crop_nhwc_bf16")


def convlstm_pointwise(gates_f32, c_f32, h_bf16):
    P, Ch = c_f32.shape
    _lib.check(_lib.load().evfly_convlstm_pointwise_nhwc(_lib.ptr(gates_f32), _lib.ptr(c_f32), h_bf16.data_ptr(), P, Ch,
                                                          _lib.stream_ptr()), "evfly_convlstm_pointwise_nhwc")


def nchw_to_grid(x_f32, Hp, Wp) -> Grid:
    N, Cc, vh, vw = x_f32.shape
    out = new_grid(N, Hp, Wp, Cc, vh, vw, x_f32.device)
    _lib.check(_lib.load().evfly_nchw_f32_to_nhwc_bf16(_lib.ptr(x_f32.contiguous()), out.data.data_ptr(), N, Cc, vh, vw, Hp, Wp,
                                                        _lib.stream_ptr()), "evfly_nchw_f32_to_nhwc_bf16")
    return out


def grid_to_nchw(data, vh, vw) -> torch.Tensor:
    """bf16 or fp32 [N,Hp,Wp,C] -> fp32 NCHW [N,C,vh,vw]."""
    N, Hp, Wp, Cc = data.shape
    out = torch.empty((N, Cc, vh, vw), dtype=torch.float32, device=data.device)
    _lib.check(_lib.load().evfly_nhwc_to_nchw_f32(data.data_ptr(), int(data.dtype == torch.float32), _lib.ptr(out), N, Cc, vh, vw,
                                                   Hp, Wp, _lib.stream_ptr()), "evfly_nhwc_to_nchw_f32")
    return out


# ---- bf16 ViT stage helpers ------------------------------------------------------------------------
def gemm_tokens(x2d, w_packed, bias=None, res_bf16=None, relu=False):
    """bf16 [M,K] @ [N,K]^T + bias (+ bf16 residual) -> bf16 [M,N] on the tensor cores."""
    M, K = x2d.shape
    Nn = w_packed.shape[0]
    assert x2d.is_contiguous() and x2d.dtype == BF16
    out = torch.empty((M, Nn), dtype=BF16, device=x2d.device)
    a = _lib.TcConvArgs()
    a.x, a.w, a.bias, a.out = x2d.data_ptr(), w_packed.data_ptr(), _lib.ptr(bias), out.data_ptr()
    a.res_bf16 = None if res_bf16 is None else res_bf16.data_ptr()
    a.M_rows, a.out_ld = M, Nn
    a.Cin, a.n_rows, a.taps, a.w_pitch, a.relu, a.out_c0 = K, Nn, 1, 0, int(relu), 0
    _call(a)
    return out


def pack_conv_kc(w: torch.Tensor) -> torch.Tensor:
    """[Cout,Cin,k,k] fp32 -> fp32 [k*k*Cin, Cout] with K index (kh*k+kw)*Cin + ci."""
    return w.permute(2, 3, 1, 0).reshape(-1, w.shape[0]).contiguous()


def patch_embed_ln(x, x_is_f32_nchw, w_kc, bias, gamma, beta, B, H, W, Cin, Cout, k, stride, pad, eps):
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    tok = torch.empty((B, OH * OW, Cout), dtype=BF16, device=x.device)
    _lib.check(_lib.load().evfly_patch_embed_ln_bf16(x.data_ptr(), int(x_is_f32_nchw), _lib.ptr(w_kc), _lib.ptr(bias), _lib.ptr(gamma),
                                                     _lib.ptr(beta), tok.data_ptr(), B, H, W, Cin, Cout, k, stride, pad, eps,
                                                     _lib.stream_ptr()), "evfly_patch_embed_ln_bf16")
    return tok, OH, OW


def layernorm_bf16(x, gamma, beta, eps):
    y = torch.empty_like(x)
    Cc = x.shape[-1]
    _lib.check(_lib.load().evfly_layernorm_bf16(x.data_ptr(), _lib.ptr(gamma), _lib.ptr(beta), y.data_ptr(), x.numel() // Cc, Cc, eps,
                                                _lib.stream_ptr()), "evfly_layernorm_bf16")
    return y


def attention_small_bf16(q, kv, heads):
    B, N, Cc = q.shape
    out = torch.empty_like(q)
    _lib.check(_lib.load().evfly_attention_small_bf16(q.data_ptr(), kv.data_ptr(), out.data_ptr(), B, N, Cc, heads, kv.shape[1],
                                                      _lib.stream_ptr()), "evfly_attention_small_bf16")
    return out


def dwconv3x3_gelu(x_bhwc, w, bias):
    B, H, W, Ce = x_bhwc.shape
    y = torch.empty_like(x_bhwc)
    _lib.check(_lib.load().evfly_dwconv3x3_gelu_nhwc_bf16(x_bhwc.data_ptr(), _lib.ptr(w), _lib.ptr(bias), y.data_ptr(), B, H, W, Ce,
                                                          _lib.stream_ptr()), "evfly_dwconv3x3_gelu_nhwc_bf16")
    return y


def pack_lstm_whh_pairs(w_hh: torch.Tensor) -> torch.Tensor:
    """W_hh [4H,H] -> bf16 [H/2, 4H, 2] = {W[r][2j], W[r][2j+1]} at [j][r]."""
    G, H = w_hh.shape
    return w_hh.t().reshape(H // 2, 2, G).permute(0, 2, 1).to(BF16).contiguous()


def gemm_into_f32(x2d, w_packed, bias, out_buf, out_c0):
    """bf16 [M,K] @ [N,K]^T + bias -> fp32 columns [out_c0, out_c0+N) of the row-major buffer out_buf [M, ld]."""
    M, K = x2d.shape
    Nn = w_packed.shape[0]
    assert x2d.is_contiguous() and x2d.dtype == BF16 and out_buf.dtype == torch.float32 and out_buf.is_contiguous()
    a = _lib.TcConvArgs()
    a.x, a.w, a.bias, a.out_f32 = x2d.data_ptr(), w_packed.data_ptr(), _lib.ptr(bias), out_buf.data_ptr()
    a.M_rows, a.out_ld = M, out_buf.shape[1]
    a.Cin, a.n_rows, a.taps, a.w_pitch, a.relu, a.out_c0 = K, Nn, 1, 0, 0, out_c0
    _call(a)
    return out_buf


# ---- fused MixFFN block (csrc/vit_fused.cu) ---------------------------------------------------------
FUSED_FFN_MIN_BATCH = 8      # below this one CTA per sample leaves the GPU empty: the per-op path is used (batch-1 streaming)
FUSED_FFN_SHAPES = {(15, 23, 32), (8, 12, 64)}


def _kmajor_swizzled(mat: torch.Tensor) -> torch.Tensor:
    """[N, K] (K = 32 or 64) -> uint8 image [N*K*2] of a K-major UMMA operand whose base is 1024-byte aligned: row n holds
    its K bf16 values in 16-byte chunks, chunk c stored at position c ^ ((n >> 1) & 3) (64-byte rows: SWIZZLE_64B) or
    c ^ (n & 7) (128-byte rows: SWIZZLE_128B) -- what TMA would write and tcgen05.mma reads."""
    N, K = mat.shape
    assert K in (32, 64)
    m = mat.to(BF16).contiguous().view(torch.int16).view(N, K // 8, 8)
    n = torch.arange(N, device=mat.device)
    c = torch.arange(K // 8, device=mat.device)
    x = ((n >> 1) & 3) if K == 32 else (n & 7)
    dst = (c[None, :] ^ x[:, None])[:, :, None].expand(-1, -1, 8)
    out = torch.empty_like(m)
    out.scatter_(1, dst, m)
    return out.reshape(-1).view(torch.uint8)


def pack_vit_ffn(mlp1_w, mlp1_b, dw_w, dw_b, mlp2_w, mlp2_b, ln_w, ln_b):
    """Operand images of one MixFFN block for evfly_vit_ffn_bf16: per 32-channel slice of the expanded activation the
    mlp1 rows [32, C], the nine conv taps as 32x32 block-diagonal matrices of the slice's four 8x8 groups
    (depthwise.weight is [8C, 8, 3, 3] with groups = C: output channel co reads input channels 8*(co//8) .. +8), and the
    mlp2 columns [C, 32]; all K-major and pre-swizzled. Returns (uint8 image, fp32 [8C + 8C + 3C] biases + LayerNorm)."""
    Ce, Cc = mlp1_w.shape
    dev = mlp1_w.device
    parts = []
    for j in range(Ce // 32):
        sl = slice(32 * j, 32 * j + 32)
        parts.append(_kmajor_swizzled(mlp1_w[sl]))
        wj = dw_w[sl]                                        # [32, 8, 3, 3]
        for tap in range(9):
            bt = torch.zeros((32, 32), dtype=torch.float32, device=dev)
            for gq in range(4):
                bt[8 * gq:8 * gq + 8, 8 * gq:8 * gq + 8] = wj[8 * gq:8 * gq + 8, :, tap // 3, tap % 3]
            parts.append(_kmajor_swizzled(bt))
        parts.append(_kmajor_swizzled(mlp2_w[:, sl]))
    img = torch.cat(parts).contiguous()
    need = _lib.load().evfly_vit_ffn_image_bytes(Cc)
    assert img.numel() == need, (img.numel(), need)
    fb = torch.cat([mlp1_b, dw_b, mlp2_b, ln_w, ln_b]).to(torch.float32).contiguous()
    return img, fb


def vit_ffn(tok, img, fbias, B, H, W, eps):
    """tok bf16 [B, H*W, C] -> LayerNorm(tok + MixFFN(tok)) bf16 [B, H*W, C] in one launch."""
    Cc = tok.shape[-1]
    assert tok.is_contiguous() and tok.dtype == BF16
    out = torch.empty_like(tok)
    _lib.check(_lib.load().evfly_vit_ffn_bf16(tok.data_ptr(), img.data_ptr(), _lib.ptr(fbias), out.data_ptr(), B, H, W, Cc, float(eps),
                                              _lib.stream_ptr()), "evfly_vit_ffn_bf16")
    return out


FUSED_ATTN = True


def pack_vit_attn(q_w, q_b, f_w, f_b):
    """query / finalLayer weights [C, C] as pre-swizzled K-major operand images + their biases (fp32 [2C])."""
    img = torch.cat([_kmajor_swizzled(q_w), _kmajor_swizzled(f_w)]).contiguous()
    return img, torch.cat([q_b, f_b]).to(torch.float32).contiguous()


def vit_attn(tok, kv, img, bias, heads):
    """tok bf16 [B,N,C], kv bf16 [B,n_kv,2C] -> tok + finalLayer(attention(query(tok), kv)) bf16 [B,N,C], one launch."""
    B, N, Cc = tok.shape
    assert tok.is_contiguous() and kv.is_contiguous() and tok.dtype == BF16 and kv.dtype == BF16
    out = torch.empty_like(tok)
    _lib.check(_lib.load().evfly_vit_attn_bf16(tok.data_ptr(), kv.data_ptr(), img.data_ptr(), _lib.ptr(bias), out.data_ptr(), B, N, Cc, heads,
                                               kv.shape[1], _lib.stream_ptr()), "evfly_vit_attn_bf16")
    return out


USE_STAGE_ABI = True    # OrigUNet in the shipped configuration through ONE C call (evfly_unet_forward, csrc/stages.cu)
FUSE_STEM = True     # binary-input stem as a table lookup inside the e12 kernel (form_BEV = 2)


def _call_stem_e12(*args):
    """13 arguments: evfly_tc_stem_e12_pool_bf16; a 14th (skip_OH): the _rows variant that only writes the y_e1 rows the skip reads."""
    if len(args) == 14:
        _lib.check(_lib.load().evfly_tc_stem_e12_pool_rows_bf16(*args, _lib.stream_ptr()), "evfly_tc_stem_e12_pool_rows_bf16")
    else:
        _lib.check(_lib.load().evfly_tc_stem_e12_pool_bf16(*args, _lib.stream_ptr()), "evfly_tc_stem_e12_pool_bf16")


def stem_e12_pool(mask_f32, stem_w, stem_b, w_packed, bias, relu=True, frames_cutoff=None, skip_rows=None):
    """Binary mask [N,1,H,W] -> (y_e1 grid, pooled grid): unet_e11 + unet_e12 + MaxPool2d(2) without e11 in HBM.
    frames_cutoff = c: the first argument is the NORMALISED FRAME instead; form_input's cutoff (in place, like the reference)
    and the mask are folded into the pattern extraction (one pass, no mask tensor).
    skip_rows = OH: y_e1 will only be read by resize_bilinear_into(..., OH, ...) (the 'interp' skip): the rows that resize does
    not sample are not written (undefined content)."""
    N, _, H, W = mask_f32.shape
    dev = mask_f32.device
    pat = torch.empty((N, H - 2, W - 2), dtype=torch.int16, device=dev)
    if frames_cutoff is None:
        _lib.check(_lib.load().evfly_stem_patterns(_lib.ptr(mask_f32), pat.data_ptr(), N, H, W, _lib.stream_ptr()), "evfly_stem_patterns")
    else:
        _lib.check(_lib.load().evfly_form_patterns(_lib.ptr(mask_f32), float(frames_cutoff), pat.data_ptr(), N, H, W, _lib.stream_ptr()), "evfly_form_patterns")
    out = new_grid(N, H, W, 32, H - 4, W - 4, dev)
    ph, pw = (H - 4) // 2, (W - 4) // 2
    pooled = new_grid(N, ph, pw, 32, ph, pw, dev)
    args = (pat.data_ptr(), _lib.ptr(stem_w.contiguous()), _lib.ptr(stem_b), w_packed.data_ptr(), _lib.ptr(bias), out.data.data_ptr(),
            pooled.data.data_ptr(), N, H, W, int(relu), pooled.Hp, pooled.Wp)
    _call_stem_e12(*args, int(skip_rows)) if (skip_rows and SKIP_ROWS) else _call_stem_e12(*args)
    return out, pooled


def conv3x3_out1(g: Grid, w_packed, bias, w1_f32, b1, relu=True) -> torch.Tensor:
    """3x3 valid conv (+bias, ReLU) followed by a 1x1 conv to one channel in the epilogue (unet_d42 + unet_out): fp32
    [N, Hp, Wp] on g's pitch grid, valid (vh-2) x (vw-2); the Cout-channel activation is never written."""
    Cout = w_packed.shape[0]
    out = torch.empty((g.N, g.Hp, g.Wp), dtype=torch.float32, device=g.data.device)
    _call_halo_out1(g.data.data_ptr(), w_packed.data_ptr(), _lib.ptr(bias), _lib.ptr(w1_f32), _lib.ptr(b1), out.data_ptr(), g.N, g.Hp, g.Wp,
                    g.vh, g.vw, g.C, Cout, int(relu))
    return out


def _call_halo_out1(*args):
    _lib.check(_lib.load().evfly_tc_conv3x3_halo_out1_bf16(*args, _lib.stream_ptr()), "evfly_tc_conv3x3_halo_out1_bf16")
