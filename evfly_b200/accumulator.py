"""Drop-in for the two C++ ROS accumulator nodes (evfly_ros/src/node.cpp and
evfly_dv_ros/src/node.cpp): a 128-biased uint8 event image, ++/-- per event, published and
reset at 30 Hz. The per-event loop runs as a scatter kernel on the B200; the class keeps the
reference's member names so that the ROS shell around it reads the same.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .events import EVENT_DTYPE, L1, _device, to_device


class ImagePublisher:
    """`saturate=False` reproduces evfly_ros (uint8 wraps, node.cpp:33-37); `saturate=True`
    reproduces evfly_dv_ros (clamped at 0/255, node.cpp:33-41), exactly, for any event order."""

    IMAGE_WIDTH = 640
    IMAGE_HEIGHT = 480
    PUBLISH_RATE = 30

    def __init__(self, saturate: bool = False, width: int | None = None, height: int | None = None,
                 device=None, staging_events: int = 1 << 20):
        _lib.load()
        self.IMAGE_WIDTH = int(width or self.IMAGE_WIDTH)
        self.IMAGE_HEIGHT = int(height or self.IMAGE_HEIGHT)
        self.saturate = bool(saturate)
        self.device = _device(device)
        self._counts = torch.zeros((2, self.IMAGE_HEIGHT, self.IMAGE_WIDTH), dtype=torch.int32,
                                   device=self.device)
        self._pinned = torch.empty((staging_events, 16), dtype=torch.uint8, pin_memory=True)
        self._window_events: list[torch.Tensor] = []  # kept for the exact replay (saturate only)
        self.image_data_ = np.full(self.IMAGE_WIDTH * self.IMAGE_HEIGHT, 128, dtype=np.uint8)

    # node.cpp:24-40 -- one EventArray message
    def eventArrayCallback(self, events) -> None:
        """`events`: EVENT_DTYPE numpy array / raw bytes of msg->events / uint8 tensor [n,16]."""
        rec = to_device(events, self.device, self._pinned)
        if rec.shape[0] == 0:
            return
        L1.accumulate_counts(rec, self.IMAGE_HEIGHT, self.IMAGE_WIDTH, out=self._counts)
        if self.saturate:
            self._window_events.append(rec)
        # the pinned staging buffer is reused by the next callback: wait for the H2D copy
        torch.cuda.current_stream().synchronize()

    # node.cpp:42-59 -- publish the image and reset it to 128
    def timerCallback(self) -> np.ndarray:
        H, W = self.IMAGE_HEIGHT, self.IMAGE_WIDTH
        mode = _lib.U8_SATURATE if self.saturate else _lib.U8_WRAP
        frame, flagged, nfl = L1.counts_to_u8(self._counts, mode, None)
        if self.saturate:
            n_flagged = int(nfl.item())
            if n_flagged > flagged.shape[0]:  # more order-dependent pixels than the list holds
                frame, flagged, nfl = L1.counts_to_u8(self._counts, mode, None, flagged_cap=n_flagged)
            if n_flagged > 0:
                rec = torch.cat(self._window_events) if len(self._window_events) > 1 else self._window_events[0]
                L1.u8_saturate_replay(rec, H, W, None, flagged, n_flagged, frame)
            self._window_events.clear()
        self.image_data_ = frame.reshape(-1).cpu().numpy()
        self._counts.zero_()
        return self.image_data_
