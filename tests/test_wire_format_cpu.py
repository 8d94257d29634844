"""CPU: the host packer of the 4-byte wire record (evfly_b200.events.pack_ev4_host) decoded again in numpy -- every event of
every window comes back with its coordinates, polarity and exact time; escape records appear where gaps exceed 4095 us; the
per-chunk time table is what the kernel adds to its running sums; streams off the 1 us grid are refused (the 8-byte format
takes them)."""
import numpy as np
import pytest

from evfly_b200 import _lib
from evfly_b200.events import CHUNK_EVENTS, EV4_SKIP, pack_ev4_host, records_time_ns
from evfly_b200.synthetic import synthetic_stream


def test_chunk_size_constant_matches_the_library():
    assert _lib.load().evfly_accumulate_chunk_events() == CHUNK_EVENTS


@pytest.mark.parametrize("T,n_per,dur_us", [(6, 100_000, 33_333), (5, 6, 33_333), (2, 300, 2_000_000), (3, 0, 1_000)])
def test_pack_ev4_round_trip(T, n_per, dur_us):
    rec, edges = synthetic_stream(5, T, n_per, 260, 346, dur_ns=dur_us * 1000, grid_ns=1000)
    edges = edges.copy()
    edges[0] += 7_000
    r4, offs, cb = pack_ev4_host(rec, edges)
    t = records_time_ns(rec)
    ci = 0
    for w in range(T):
        seg = r4[offs[w]:offs[w + 1]]
        dt = np.cumsum((seg >> 20).astype(np.int64))
        real = (seg & 0x7FFFF) != EV4_SKIP
        sel = (t >= edges[w]) & (t < edges[w + 1])
        assert real.sum() == sel.sum()
        assert np.array_equal(dt[real] * 1000, t[sel] - edges[w])
        assert np.array_equal(seg[real] & 1023, rec["x"][sel]) and np.array_equal((seg[real] >> 10) & 511, rec["y"][sel])
        assert np.array_equal((seg[real] >> 19) & 1, rec["polarity"][sel])
        for k in range(-(-len(seg) // CHUNK_EVENTS)):
            assert cb[ci] == (0 if k == 0 else dt[k * CHUNK_EVENTS - 1])
            ci += 1
    assert ci == len(cb)
    if 0 < n_per <= 300:
        assert ((r4 & 0x7FFFF) == EV4_SKIP).any()


def test_pack_ev4_refuses_what_it_cannot_represent():
    rec, edges = synthetic_stream(1, 3, 1000, 260, 346)                       # nanosecond timestamps
    assert pack_ev4_host(rec, edges) is None
    rec, edges = synthetic_stream(1, 3, 1000, 260, 346, dur_ns=2_000_000, grid_ns=1000)
    assert pack_ev4_host(rec[::-1].copy(), edges) is None                      # not sorted by time
    assert pack_ev4_host(rec, edges + 1) is None                               # edges off the grid
    assert pack_ev4_host(rec, edges) is not None
    big = rec.copy()
    big["x"][5] = 1100                                                        # a sensor wider than 1023 pixels
    assert pack_ev4_host(big, edges) is None
