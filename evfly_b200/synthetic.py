"""Seeded synthetic event streams of the shapes SURVEY.md 8(d) names (there is no dataset and
no network on the benchmark box). Host-side numpy only; shared by tests/ and bench.py."""
from __future__ import annotations

import numpy as np

from .events import EVENT_DTYPE, NS, make_records


def synthetic_window(seed: int, n_events: int, H: int, W: int, *, t0_ns: int = 0,
                     dur_ns: int = 33_333_333, distribution: str = "uniform", grid_ns: int = 1) -> np.ndarray:
    """One window of events sorted by time. grid_ns: timestamp resolution (1000 = the 1 us grid of a DVS sensor).

    uniform   : x~U{0..W-1}, y~U{0..H-1}, p~Bern(.5), t sorted U[t0, t0+dur)   (cfg 1 / 2)
    clustered : 80 % of the events on 2 % of the pixels (edge-like hot spots)   (cfg 2)
    """
    rng = np.random.default_rng(seed)
    if distribution == "uniform":
        x = rng.integers(0, W, n_events, dtype=np.int64)
        y = rng.integers(0, H, n_events, dtype=np.int64)
    elif distribution == "clustered":
        n_hot_px = max(1, int(0.02 * H * W))
        hot = rng.choice(H * W, n_hot_px, replace=False)
        is_hot = rng.random(n_events) < 0.8
        pix = np.where(is_hot, hot[rng.integers(0, n_hot_px, n_events)],
                       rng.integers(0, H * W, n_events))
        x, y = pix % W, pix // W
    else:
        raise ValueError(distribution)
    p = rng.integers(0, 2, n_events, dtype=np.int64)
    t = np.sort(rng.integers(0, dur_ns // grid_ns, n_events, dtype=np.int64)) * grid_ns + t0_ns
    return make_records(x, y, t, p)


def synthetic_stream(seed: int, n_windows: int, events_per_window: int, H: int, W: int, *,
                     dur_ns: int = 33_333_333, distribution: str = "uniform", grid_ns: int = 1):
    """`n_windows` consecutive windows of one stream. Returns (records, edges_ns int64 [T+1])."""
    parts = [synthetic_window(seed * 100003 + w, events_per_window, H, W, t0_ns=w * dur_ns,
                              dur_ns=dur_ns, distribution=distribution, grid_ns=grid_ns) for w in range(n_windows)]
    edges = np.arange(n_windows + 1, dtype=np.int64) * dur_ns
    return np.concatenate(parts), edges


def records_to_rows(rec: np.ndarray, pol_neg_value: float = 0.0) -> np.ndarray:
    """EVENT_DTYPE -> the float64 [n,4] = (t_ns, x, y, p) rows form_eventframe is called with."""
    rows = np.empty((rec.shape[0], 4), dtype=np.float64)
    rows[:, 0] = rec["ts_sec"].astype(np.float64) * NS + rec["ts_nsec"]
    rows[:, 1] = rec["x"]
    rows[:, 2] = rec["y"]
    rows[:, 3] = np.where(rec["polarity"] == 1, 1.0, pol_neg_value)
    return rows
