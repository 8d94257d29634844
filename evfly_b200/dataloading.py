"""Drop-in for the event-frame normalisation inside evfly's learner/dataloading.py (:508-533) -- the step
between accumulation and the model on the offline path ("next" row N1). File IO stays with the reference.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .events import _device


def normalize_event_frames(ev, rescale_evs: float = -1.0, evs_min_cutoff=None) -> torch.Tensor:
    """One trajectory of event frames [T,H,W] -> normalised frames (float32 CUDA tensor), exactly as
    dataloading.py does per trajectory:
      rescale_evs > 0   : clamp(ev / rescale_evs, -1, 1)                                   (:511-512)
      rescale_evs == -1 : per frame s = quantile(|ev|, 0.97); clamp(ev / s, -1, 1)         (:515-524)
      evs_min_cutoff    : |ev| < cutoff -> 0                                               (:531-533)
    """
    lib = _lib.load()
    dev = _device()
    x = torch.as_tensor(ev).to(device=dev, dtype=torch.float32).contiguous()
    T = x.shape[0]
    elems = x[0].numel()
    cutoff = 0.0 if evs_min_cutoff is None else float(evs_min_cutoff)
    out = torch.empty_like(x)
    if rescale_evs > 0.0:
        ops.map4d(x.view(T, -1), out.view(T, -1), div=float(rescale_evs), lo=-1.0, hi=1.0)
        if cutoff > 0:
            _lib.check(lib.evfly_min_cutoff_f32(_lib.ptr(out), out.numel(), cutoff, _lib.stream_ptr()), "evfly_min_cutoff_f32")
    elif rescale_evs == -1.0:
        _lib.check(lib.evfly_quantile_scale_clip(_lib.ptr(x), T, elems, 0.97, -1.0, 1.0, cutoff, _lib.ptr(out), None,
                                                 _lib.stream_ptr()), "evfly_quantile_scale_clip")
    else:
        out.copy_(x)
        if cutoff > 0:
            _lib.check(lib.evfly_min_cutoff_f32(_lib.ptr(out), out.numel(), cutoff, _lib.stream_ptr()), "evfly_min_cutoff_f32")
    return out


def resize_trajectory(x, resize_input) -> torch.Tensor:
    """dataloading.py:401-416: F.interpolate(traj.unsqueeze(1), size=resize_input, mode='bilinear',
    align_corners=False).squeeze() for one trajectory of images / depths / event frames [T,H,W] -> [T,h,w]
    (float32 CUDA tensor). Same sampling positions and lerp order as torch (evfly_resize_bilinear_f32)."""
    dev = _device()
    x = torch.as_tensor(x).to(device=dev, dtype=torch.float32).contiguous()
    T, H, W = x.shape
    h, w = resize_input
    if (H, W) == (h, w):
        return x
    return ops.resize_bilinear(x.view(T, 1, H, W), (h, w), align_corners=False).view(T, h, w)
