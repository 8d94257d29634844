"""Drop-in for the difflog event approximation of the simulator front end ("next" row N2):
envtest/ros/run_competition.py::compute_events (:603-635) and utils/to_events.py:417-439."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .events import _device

SMALL_EPS = 1e-5      # run_competition.py:29


def compute_events(im, prev_im, neg_thresh=0.2, pos_thresh=0.2, *, inputs_are_log=False, shape=None):
    """Estimated events frame from two images (float64, like the reference's numpy code).
    `im is None or prev_im is None` -> zeros of `shape` (run_competition.py:616-618)."""
    if im is None or prev_im is None:
        return np.zeros(shape)
    lib = _lib.load()
    dev = _device()
    a = torch.as_tensor(np.asarray(im, dtype=np.float64)).to(dev).contiguous()
    b = torch.as_tensor(np.asarray(prev_im, dtype=np.float64)).to(dev).contiguous()
    out = torch.empty_like(a)
    ws = torch.empty((1,), dtype=torch.int64, device=dev)
    _lib.check(lib.evfly_difflog_events_f64(_lib.ptr(a), _lib.ptr(b), a.numel(), SMALL_EPS, int(inputs_are_log), float(pos_thresh),
                                            float(neg_thresh), _lib.ptr(out), _lib.ptr(ws), _lib.stream_ptr()), "evfly_difflog_events_f64")
    return out.cpu().numpy()
