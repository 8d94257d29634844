"""Typed Python wrappers of the L3 operator entry points (include/evfly_b200.h, model section).

Tensors are fp32 CUDA tensors. Views are welcome wherever a `*_view` tensor is named: only
data_ptr() and stride() are passed down, so permute / reshape-views / channel slices of a concat
buffer cost nothing. Nothing here computes on the host and nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import ACT

INF = float("inf")


def _f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda:
        raise _lib.EvflyError(f"{what}: expected a float32 CUDA tensor, got {t.dtype} on {t.device}")
    return t


def conv2d(x_view, w, bias=None, *, stride=1, pad=0, groups=1, act=None, post=None, res_view=None,
           out_view=None):
    """y = post_scale*act(conv(x)+bias)+post_shift (+res). x_view: 4-D (any strides) [N,Cin,H,W];
    w: contiguous [Cout,Cin/groups,KH,KW]; out_view: 4-D view to write (allocated NCHW if None);
    res_view must alias the layout of out_view."""
    lib = _lib.load()
    _f32(x_view, "conv2d x"); _f32(w, "conv2d w")
    N, Cin, H, W = x_view.shape
    Cout, _, KH, KW = w.shape
    OH = (H + 2 * pad - KH) // stride + 1
    OW = (W + 2 * pad - KW) // stride + 1
    if out_view is None:
        out_view = torch.empty((N, Cout, OH, OW), dtype=torch.float32, device=x_view.device)
    assert tuple(out_view.shape) == (N, Cout, OH, OW), (tuple(out_view.shape), (N, Cout, OH, OW))
    if res_view is not None:  # same addressing as the output (strides of size-1 dims are irrelevant)
        assert res_view.shape == out_view.shape
        assert all(d == 1 or a == b for d, a, b in zip(out_view.shape, res_view.stride(), out_view.stride()))
    a = _lib.ConvArgs()
    a.x, a.w, a.y = x_view.data_ptr(), _lib.ptr(w), out_view.data_ptr()
    a.bias = _lib.ptr(bias)
    a.post_scale, a.post_shift = (None, None) if post is None else (_lib.ptr(post[0]), _lib.ptr(post[1]))
    a.res = None if res_view is None else res_view.data_ptr()
    a.N, a.Cin, a.H, a.W, a.Cout, a.KH, a.KW = N, Cin, H, W, Cout, KH, KW
    a.stride, a.pad, a.groups, a.act = stride, pad, groups, ACT[act]
    a.xs = (C.c_int64 * 4)(*x_view.stride())
    a.ys = (C.c_int64 * 4)(*out_view.stride())
    _lib.check(lib.evfly_conv2d_f32(C.byref(a), _lib.stream_ptr()), "evfly_conv2d_f32")
    return out_view


def linear(x2d, w, bias=None, *, act=None, res2d=None, out2d=None):
    """y[M,N] = act(x[M,K] @ w[N,K]^T + bias) (+res): the 1x1 case of conv2d. Rows may be strided."""
    M, K = x2d.shape
    Nout = w.shape[0]
    if out2d is None:
        out2d = torch.empty((M, Nout), dtype=torch.float32, device=x2d.device)
    if M <= 8 and x2d.stride(1) == 1 and out2d.stride(1) == 1 and (res2d is None or res2d.stride(1) == 1) and M * K * 4 <= 160 * 1024:
        _lib.check(_lib.load().evfly_linear_smallm_f32(
            x2d.data_ptr(), x2d.stride(0), _lib.ptr(w), _lib.ptr(bias), None if res2d is None else res2d.data_ptr(),
            0 if res2d is None else res2d.stride(0), out2d.data_ptr(), out2d.stride(0), M, Nout, K, ACT[act], _lib.stream_ptr()),
            "evfly_linear_smallm_f32")
        return out2d
    as4 = lambda t: t.as_strided((1, t.shape[1], 1, t.shape[0]), (0, t.stride(1), 0, t.stride(0)), t.storage_offset())
    conv2d(as4(x2d), w.reshape(Nout, K, 1, 1), bias, act=act,
           res_view=None if res2d is None else as4(res2d), out_view=as4(out2d))
    return out2d


def pool2d(x, k, stride, mode="max", negate_in=False, negate_out=False):
    lib = _lib.load()
    x = _f32(x, "pool2d").contiguous()
    N, Cc, H, W = x.shape
    OH, OW = (H - k) // stride + 1, (W - k) // stride + 1
    y = torch.empty((N, Cc, OH, OW), dtype=torch.float32, device=x.device)
    _lib.check(lib.evfly_pool2d_f32(_lib.ptr(x), _lib.ptr(y), N * Cc, H, W, k, stride, 0 if mode == "max" else 1,
                                    int(negate_in), int(negate_out), _lib.stream_ptr()), "evfly_pool2d_f32")
    return y


def _s4(t):
    return (C.c_int64 * 4)(*t.stride())


def resize_bilinear(x_view, size, align_corners=False, out_view=None, mul=1.0, add=0.0, lo=-INF, hi=INF, pre=None):
    """pre=(mul, lo, hi): interpolate clip(x * mul, lo, hi) instead of x (fused, nothing materialised)."""
    lib = _lib.load()
    _f32(x_view, "resize_bilinear")
    N, Cc, H, W = x_view.shape
    OH, OW = size
    if out_view is None:
        out_view = torch.empty((N, Cc, OH, OW), dtype=torch.float32, device=x_view.device)
    assert tuple(out_view.shape) == (N, Cc, OH, OW)
    if pre is not None:
        assert (mul, add, lo, hi) == (1.0, 0.0, -INF, INF)
        _lib.check(lib.evfly_resize_bilinear_premap_f32(x_view.data_ptr(), _s4(x_view), out_view.data_ptr(), _s4(out_view), N, Cc, H, W,
                                                        OH, OW, int(align_corners), float(pre[0]), float(pre[1]), float(pre[2]),
                                                        _lib.stream_ptr()), "evfly_resize_bilinear_premap_f32")
        return out_view
    _lib.check(lib.evfly_resize_bilinear_f32(x_view.data_ptr(), _s4(x_view), out_view.data_ptr(), _s4(out_view), N, Cc, H, W,
                                             OH, OW, int(align_corners), mul, add, lo, hi, _lib.stream_ptr()),
               "evfly_resize_bilinear_f32")
    return out_view


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    lib = _lib.load()
    x = _f32(x, "layernorm")
    assert x.is_contiguous()
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    if out is None:
        out = torch.empty_like(x)
    _lib.check(lib.evfly_layernorm_f32(_lib.ptr(x), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(out), rows, Cc, eps,
                                       _lib.stream_ptr()), "evfly_layernorm_f32")
    return out


def attention_small(q, kv, heads):
    lib = _lib.load()
    B, N, Cc = q.shape
    n_kv = kv.shape[1]
    assert q.is_contiguous() and kv.is_contiguous() and kv.shape[2] == 2 * Cc
    out = torch.empty_like(q)
    _lib.check(lib.evfly_attention_small_f32(_lib.ptr(q), _lib.ptr(kv), _lib.ptr(out), B, N, Cc, heads, n_kv,
                                             _lib.stream_ptr()), "evfly_attention_small_f32")
    return out


def map4d(x_view, out_view=None, mul=1.0, div=1.0, add=0.0, lo=-INF, hi=INF):
    """out = clip(x*mul/div + add, lo, hi) on (up to) 4-D views of any strides."""
    lib = _lib.load()
    _f32(x_view, "map4d")
    if out_view is None:
        out_view = torch.empty(x_view.shape, dtype=torch.float32, device=x_view.device)
    assert x_view.shape == out_view.shape and x_view.dim() <= 4
    pad = 4 - x_view.dim()
    dims = (C.c_int64 * 4)(*([1] * pad + list(x_view.shape)))
    xs = (C.c_int64 * 4)(*([0] * pad + list(x_view.stride())))
    ys = (C.c_int64 * 4)(*([0] * pad + list(out_view.stride())))
    _lib.check(lib.evfly_map4d_f32(x_view.data_ptr(), xs, out_view.data_ptr(), ys, dims, mul, div, add, lo, hi,
                                   _lib.stream_ptr()), "evfly_map4d_f32")
    return out_view


def pixel_shuffle(x_view, r, out_view):
    lib = _lib.load()
    _f32(x_view, "pixel_shuffle")
    N, Cr, H, W = x_view.shape
    Cc = Cr // (r * r)
    assert tuple(out_view.shape) == (N, Cc, H * r, W * r)
    _lib.check(lib.evfly_pixel_shuffle_f32(x_view.data_ptr(), _s4(x_view), out_view.data_ptr(), _s4(out_view), N, Cc, H, W, r,
                                           _lib.stream_ptr()), "evfly_pixel_shuffle_f32")
    return out_view


def form_input(frames, form_bev, cutoff):
    """In place on `frames` (like the reference) + returns the UNet input."""
    lib = _lib.load()
    _f32(frames, "form_input")
    assert frames.is_contiguous() and frames.shape[1] == 1
    N, _, H, W = frames.shape
    out = torch.empty((N, 2 if form_bev == 0 else 1, H, W), dtype=torch.float32, device=frames.device)
    _lib.check(lib.evfly_form_input_f32(_lib.ptr(frames), _lib.ptr(out), frames.numel(), H * W, form_bev, cutoff,
                                        _lib.stream_ptr()), "evfly_form_input_f32")
    return out


def lstm_layer_seq(gx, whh_t, h0, c0):
    """gx [T,4H] (input projection incl. biases), whh_t [H,4H]. Returns (hs [T,H], hT [H], cT [H])."""
    lib = _lib.load()
    T, G = gx.shape
    H = G // 4
    hs = torch.empty((T, H), dtype=torch.float32, device=gx.device)
    hT = torch.empty((H,), dtype=torch.float32, device=gx.device)
    cT = torch.empty((H,), dtype=torch.float32, device=gx.device)
    _lib.check(lib.evfly_lstm_seq_f32(_lib.ptr(gx), _lib.ptr(whh_t), _lib.ptr(h0), _lib.ptr(c0), _lib.ptr(hs), _lib.ptr(hT),
                                      _lib.ptr(cT), T, H, 1, _lib.stream_ptr()), "evfly_lstm_seq_f32")
    return hs, hT, cT


def convlstm_pointwise(gates, c, h_out):
    lib = _lib.load()
    Ch = c.shape[-3]
    P = c.shape[-2] * c.shape[-1]
    _lib.check(lib.evfly_convlstm_pointwise_f32(_lib.ptr(gates), _lib.ptr(c), h_out.data_ptr(), Ch, P, _lib.stream_ptr()),
               "evfly_convlstm_pointwise_f32")
    return h_out


def velpred_unit(y):
    lib = _lib.load()
    N = y.shape[0]
    out = torch.empty((N, 3), dtype=torch.float32, device=y.device)
    _lib.check(lib.evfly_velpred_unit_f32(_lib.ptr(y.contiguous()), _lib.ptr(out), N, _lib.stream_ptr()), "evfly_velpred_unit_f32")
    return out
