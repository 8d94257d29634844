"""ctypes binding of libevfly_b200.so (the C ABI declared in include/evfly_b200.h).

There is deliberately no fallback: if the CUDA library is missing, importing a product path
raises. Nothing in this package computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import re

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libevfly_b200.so")
HEADER_PATH = os.path.join(PKG_DIR, "..", "include", "evfly_b200.h")

# ---- constants mirrored from the header ---------------------------------------------------
POL_NEG, POL_POS, POL_SKIP = 0, 1, 2
NEG_IS_ZERO, NEG_IS_NEGATIVE = 0, 1
U8_WRAP, U8_SATURATE = 0, 1


class EvflyError(RuntimeError):
    pass


_vp, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); every symbol include/evfly_b200.h declares must be listed here
# (tests/test_abi.py cross-checks this table against the header).
SIGNATURES = {
    "evfly_abi_version": (_i32, []),
    "evfly_last_error": (C.c_char_p, []),
    "evfly_launch_count": (_i64, []),
    "evfly_pack_events_f64": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _f64, _f64, _i64, _vp, _vp, _vp, _vp]),
    "evfly_pack_events_soa": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "evfly_accumulate_counts": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "evfly_accumulate_counts_binned_workspace_bytes": (_i64, [_i64, _i32, _i32]),
    "evfly_accumulate_counts_binned": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "evfly_counts_to_frame_f64": (_i32, [_vp, _i32, _i32, _f64, _f64, _vp, _vp]),
    "evfly_counts_to_u8": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "evfly_u8_saturate_replay": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "evfly_voxel_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "evfly_voxelize_window": (_i32, [_vp, _i64, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _vp, _i32, _vp]),
    "evfly_accumulate_windows": (_i32, [_vp, _i64, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "evfly_pack_events_ev8": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp, _vp]),
    "evfly_accumulate_sorted_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32, _i32]),
    "evfly_accumulate_windows_sorted": (_i32, [_vp, _i64, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _i64, _vp]),
    "evfly_accumulate_windows_ev8": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "evfly_accumulate_windows_ev4": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "evfly_accumulate_chunk_events": (_i32, []),
    "evfly_decode_crop": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp]),
    "evfly_difflog_events_f64": (_i32, [_vp, _vp, _i64, _f64, _i32, _f64, _f64, _vp, _vp, _vp]),
    "evfly_min_cutoff_f32": (_i32, [_vp, _i64, _f32, _vp]),
    "evfly_counts_normalise": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _f32, _f32, _vp, _vp, _vp]),
    "evfly_remap_bicubic_f32": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "evfly_remap_events_f32": (_i32, [_vp, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "evfly_quantile_scale_clip": (_i32, [_vp, _i32, _i64, _f32, _f32, _f32, _f32, _vp, _vp, _vp]),
}



class ConvArgs(C.Structure):
    """evfly_conv2d_args (include/evfly_b200.h)."""
    _fields_ = [("x", _vp), ("w", _vp), ("bias", _vp), ("post_scale", _vp), ("post_shift", _vp),
                ("res", _vp), ("y", _vp),
                ("N", _i32), ("Cin", _i32), ("H", _i32), ("W", _i32), ("Cout", _i32), ("KH", _i32),
                ("KW", _i32), ("stride", _i32), ("pad", _i32), ("groups", _i32), ("act", _i32),
                ("reserved", _i32),
                ("xs", _i64 * 4), ("ys", _i64 * 4)]


_p64 = C.POINTER(_i64)
SIGNATURES.update({
    "evfly_conv2d_f32": (_i32, [C.POINTER(ConvArgs), _vp]),
    "evfly_linear_smallm_f32": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "evfly_pool2d_f32": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_resize_bilinear_f32": (_i32, [_vp, _p64, _vp, _p64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _f32, _vp]),
    "evfly_resize_bilinear_premap_f32": (_i32, [_vp, _p64, _vp, _p64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _vp]),
    "evfly_layernorm_f32": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp]),
    "evfly_attention_small_f32": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_map4d_f32": (_i32, [_vp, _p64, _vp, _p64, _p64, _f32, _f32, _f32, _f32, _f32, _vp]),
    "evfly_pixel_shuffle_f32": (_i32, [_vp, _p64, _vp, _p64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_form_input_f32": (_i32, [_vp, _vp, _i64, _i64, _i32, _f32, _vp]),
    "evfly_lstm_seq_f32": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "evfly_lstm_pointwise_f32": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "evfly_convlstm_pointwise_f32": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "evfly_velpred_unit_f32": (_i32, [_vp, _vp, _i32, _vp]),
})



class TcConvArgs(C.Structure):
    """evfly_tc_conv_args (include/evfly_b200.h)."""
    _fields_ = [("x", _vp), ("w", _vp), ("bias", _vp), ("res_f32", _vp), ("out", _vp), ("out_f32", _vp),
                ("M_rows", _i64), ("out_ld", _i64),
                ("Cin", _i32), ("n_rows", _i32), ("taps", _i32), ("w_pitch", _i32), ("relu", _i32), ("out_c0", _i32),
                ("convt", _i32), ("Hp", _i32), ("Wp", _i32), ("valid_h", _i32), ("valid_w", _i32), ("cout_t", _i32),
                ("res_bf16", _vp), ("flags", _i64), ("lstm_c", _vp), ("lstm_h", _vp)]


SIGNATURES.update({
    "evfly_tc_conv_bf16": (_i32, [C.POINTER(TcConvArgs), _vp]),
    "evfly_convlstm_scan_bf16": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _vp]),
    "evfly_convlstm_scan_fused_bf16": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp]),
    "evfly_stem_conv3x3_bf16": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "evfly_stem_conv3x3_fma_bf16": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "evfly_maxpool2x2_nhwc_bf16": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_resize_bilinear_nhwc_bf16": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i64, _i32, _vp]),
    "evfly_crop_nhwc_bf16": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i64, _i32, _vp]),
    "evfly_convlstm_pointwise_nhwc": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "evfly_nchw_f32_to_nhwc_bf16": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_nhwc_to_nchw_f32": (_i32, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_patch_embed_ln_bf16": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "evfly_layernorm_bf16": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp]),
    "evfly_attention_small_bf16": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "evfly_dwconv3x3_gelu_nhwc_bf16": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "evfly_tc_conv3x3_halo_bf16": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_tc_conv3x3_halo_compact_bf16": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_tc_conv3x3_halo_out1_bf16": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_tc_conv3x3_same_bf16": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_shuffle_upsample_cat_bf16": (_i32, [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _i64, _i32, _vp]),
    "evfly_tc_conv3x3_halo_pool_bf16": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_tc_conv3x3_halo_pool_rows_bf16": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_vit_ffn_image_bytes": (_i64, [_i32]),
    "evfly_vit_ffn_bf16": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp]),
    "evfly_vit_attn_bf16": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "evfly_stem_patterns": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "evfly_form_patterns": (_i32, [_vp, _f32, _vp, _i32, _i32, _i32, _vp]),
    "evfly_tc_stem_e12_pool_bf16": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_tc_stem_e12_pool_rows_bf16": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evfly_lstm_seq_smemw": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
})

class UNetWeights(C.Structure):
    """evfly_unet_weights (include/evfly_b200.h)."""
    _fields_ = [("e11_w", _vp), ("e11_b", _vp), ("conv_w", _vp * 17), ("conv_b", _vp * 17), ("up_w", _vp * 4), ("up_b", _vp * 4),
                ("out_w", _vp), ("out_b", _vp), ("lstm_wx", _vp), ("lstm_wh", _vp)]


class VitLayerWeights(C.Structure):
    _fields_ = [(n, _vp) for n in ("red_w", "red_b", "red_ln_g", "red_ln_b", "kv_w", "kv_b", "attn_img", "attn_bias", "ffn_img", "ffn_bias")]


class VitStageWeights(C.Structure):
    _fields_ = [(n, _vp) for n in ("patch_w", "patch_b", "patch_ln_g", "patch_ln_b")] + [("layer", VitLayerWeights * 2)]


class VitLstmWeights(C.Structure):
    """evfly_vit_lstm_weights (include/evfly_b200.h)."""
    _fields_ = [("stage", VitStageWeights * 2), ("ds_w", _vp), ("ds_b", _vp), ("dec_w", _vp), ("dec_b", _vp),
                ("lstm_w_ih", _vp * 3), ("lstm_b", _vp * 3), ("lstm_whh_pairs", _vp * 3), ("lstm_whh_t", _vp * 3), ("fc2_w", _vp), ("fc2_b", _vp)]


SIGNATURES.update({
    "evfly_vit_lstm_workspace_bytes": (_i64, [_i32]),
    "evfly_vit_lstm_forward": (_i32, [C.POINTER(VitLstmWeights), _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "evfly_prep_frame": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "evfly_unet_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "evfly_unet_forward": (_i32, [C.POINTER(UNetWeights), _vp, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
})

ACT = {None: 0, "none": 0, "relu": 1, "leaky_relu": 2, "gelu": 3, "tanh": 4, "sigmoid": 5}

_lib = None


def header_symbols() -> list[str]:
    """Function names declared in include/evfly_b200.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(evfly_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    """Load the shared library (no GPU needed for loading) and set the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EvflyError(
            f"{LIB_PATH} is missing. Build it with `python -m evfly_b200._build` "
            "(__graft_entry__.build()). evfly_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.evfly_abi_version() != 1:
        raise EvflyError("libevfly_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


TC_COMPACT = 1            # EVFLY_TC_COMPACT
ERR_UNSUPPORTED = -4      # EVFLY_ERR_UNSUPPORTED (include/evfly_b200.h)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().evfly_last_error().decode(errors="replace")
        raise EvflyError(f"{what or 'libevfly_b200'} failed (rc={rc}): {msg}")


def ptr(t) -> int | None:
    """data_ptr of a CUDA tensor (None -> NULL). Refuses host tensors: the ABI is device-only."""
    if t is None:
        return None
    if not t.is_cuda:
        raise EvflyError("libevfly_b200 takes device pointers only; got a CPU tensor")
    if not t.is_contiguous():
        raise EvflyError("libevfly_b200 takes contiguous tensors only")
    return t.data_ptr()


def ptr_any(t) -> int | None:
    """data_ptr of a (possibly strided) CUDA view; the caller passes t.stride() alongside."""
    if t is None:
        return None
    if not t.is_cuda:
        raise EvflyError("libevfly_b200 takes device pointers only; got a CPU tensor")
    return t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().evfly_launch_count())
