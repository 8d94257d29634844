"""Rectification ("next" row N3): the reference's utils/calibration_tools/rectify_bag.py surface on the GPU.

`remap_img`, `remap_events` and `Aligner.align` keep the reference's names, argument meaning and return types
(rectify_bag.py:91-138); the cv2.remap(INTER_CUBIC) arithmetic runs in `evfly_remap_bicubic_f32`, bit-identical to
OpenCV. What stays on the host is the one-time construction of the maps from the calibration yaml
(`CameraSystem.getRemapping`, rectify_bag.py:57-89: cv2.initUndistortRectifyMap / undistortPoints) -- pass its
output to `Aligner(depth_map=..., davis_map=...)`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _dev_f32(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


def remap_bicubic(src: torch.Tensor, mapx: torch.Tensor, mapy: torch.Tensor, flip=False, rotate=False,
                  out: torch.Tensor | None = None) -> torch.Tensor:
    """src float32 [N,H,W] (or uint8 [N,H,W]: decoded as (v-128)*0.2 on the fly) on the device; mapx/mapy float32
    [OH,OW] (may be a row-strided window of larger maps). Returns float32 [N,OH,OW]."""
    assert src.is_cuda and src.dim() == 3 and src.is_contiguous() and src.dtype in (torch.float32, torch.uint8)
    assert mapx.shape == mapy.shape and mapx.dtype == mapy.dtype == torch.float32
    assert mapx.stride(1) == 1 and mapy.stride(1) == 1 and mapx.stride(0) == mapy.stride(0)
    N, H, W = src.shape
    OH, OW = mapx.shape
    if out is None:
        out = torch.empty((N, OH, OW), dtype=torch.float32, device=src.device)
    _lib.check(_lib.load().evfly_remap_bicubic_f32(_lib.ptr_any(src), int(src.dtype == torch.uint8), N, H, W, mapx.data_ptr(), mapy.data_ptr(),
                                                    mapx.stride(0), OH, OW, int(bool(flip)), int(bool(rotate)), _lib.ptr(out), _lib.stream_ptr()),
               "evfly_remap_bicubic_f32")
    return out


def remap_img(img, map, flip, rotate, device="cuda"):
    """rectify_bag.py:91-98. numpy in -> numpy out; a CUDA tensor in -> a CUDA tensor out."""
    mx, my = (_dev_f32(m, device) for m in map)
    as_numpy = not isinstance(img, torch.Tensor)
    src = _dev_f32(img, device)
    out = remap_bicubic(src[None], mx, my, flip=flip, rotate=rotate)[0]
    return out.cpu().numpy() if as_numpy else out


def remap_events(events, map, rotate, shape, device="cuda"):
    """rectify_bag.py:101-116: events = {'x','y','t','p'}; returns the same dict with float coordinates, masked."""
    mx, my = (_dev_f32(m, device) for m in map)
    H, W = mx.shape
    x = torch.as_tensor(np.asarray(events["x"]).astype(np.int32)).to(device)
    y = torch.as_tensor(np.asarray(events["y"]).astype(np.int32)).to(device)
    n = x.shape[0]
    ox = torch.empty((n,), dtype=torch.float32, device=device)
    oy = torch.empty((n,), dtype=torch.float32, device=device)
    keep = torch.empty((n,), dtype=torch.uint8, device=device)
    tw, th = shape
    _lib.check(_lib.load().evfly_remap_events_f32(_lib.ptr_any(x), _lib.ptr_any(y), n, mx.data_ptr(), my.data_ptr(), H, W, int(bool(rotate)),
                                                   int(tw), int(th), _lib.ptr(ox), _lib.ptr(oy), _lib.ptr_any(keep), _lib.stream_ptr()),
               "evfly_remap_events_f32")
    m = keep.bool().cpu().numpy()
    return {"x": ox.cpu().numpy()[m], "y": oy.cpu().numpy()[m], "t": np.asarray(events["t"])[m], "p": np.asarray(events["p"])[m]}


class Aligner:
    """rectify_bag.py:118-138 with the maps supplied (see the module docstring)."""

    def __init__(self, depth_map=None, davis_map=None, device="cuda"):
        self.device = torch.device(device)
        self.depth_map = None if depth_map is None else tuple(_dev_f32(m, self.device) for m in depth_map)
        self.davis_map = None if davis_map is None else tuple(_dev_f32(m, self.device) for m in davis_map)

    def align(self, depth=None, davis=None):
        out = {"depth": None, "davis": None}
        if depth is not None:
            out["depth"] = remap_img(depth, self.depth_map, flip=False, rotate=False, device=self.device)
        if davis is not None:
            out["davis"] = remap_img(davis, self.davis_map, flip=False, rotate=False, device=self.device)
        return out

    def davis_window(self, h, w):
        """The (h, w) centre window of the event-camera maps (run.py:346-351 crops AFTER aligning; computing only this
        window of the remap gives the same pixels)."""
        H, W = self.davis_map[0].shape
        r0, c0 = H // 2 - h // 2, W // 2 - w // 2
        return tuple(m[r0:r0 + h, c0:c0 + w] for m in self.davis_map)
