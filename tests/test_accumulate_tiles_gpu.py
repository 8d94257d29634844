"""GPU parity of the shared-memory-tile accumulation (csrc/accumulate_tiled.cu) and of the 8-byte wire format:
count frames bit-exact against the oracle (C restatement of np.histogram2d / node.cpp) and against the L2-reduction
kernels on the same records; voxel grids within 1e-6 of the per-cell sum of |weights| (fp32 sums in another order).
Shapes cover one chunk, many chunks, more than 1024 chunks per window (the second batch of the per-band walk),
empty windows, records outside every window, counts-only calls and time-major frame slots."""
import numpy as np
import pytest
import torch

from evfly_b200 import _lib
from evfly_b200.events import EVENT8_DTYPE, L1, WireBatch, make_records, pack_ev4_host, pack_ev8_host, to_device
from evfly_b200.synthetic import synthetic_stream, synthetic_window
from oracle import ev_oracle as O

pytestmark = pytest.mark.gpu


def _check_voxel(vox_gpu, vox_ref, vabs=None):
    err = np.abs(vox_gpu.astype(np.float64) - vox_ref)
    tol = 1e-6 * (np.maximum(1.0, vabs) if vabs is not None else np.maximum(1.0, np.abs(vox_ref)) * 8)
    assert (err <= tol).all(), float(err.max())


@pytest.mark.parametrize("H,W,T,n_per,B", [(260, 346, 7, 100_000, 5), (480, 640, 3, 100_000, 5), (120, 160, 23, 30_000, 5),
                                           (37, 53, 4, 5_000, 1), (260, 346, 2, 3_000, 9), (480, 640, 1, 700_000, 5)])
def test_tiles_equal_oracle_and_scatter(cuda_lib, H, W, T, n_per, B):
    rec, edges = synthetic_stream(11, T, n_per, H, W, dur_ns=2_000_000)
    c_ref, v_ref = O.windows(rec, edges, H, W, B=B)
    d, d_edges = to_device(rec), torch.from_numpy(edges).cuda()
    counts, vox = L1.accumulate_windows(d, d_edges, H, W, B, algo="tiles")
    assert np.array_equal(counts.cpu().numpy(), c_ref)
    _check_voxel(vox.cpu().numpy(), v_ref)
    c2, v2 = L1.accumulate_windows(d, d_edges, H, W, B, algo="scatter")
    assert torch.equal(counts, c2)
    assert (vox - v2).abs().max().item() <= 1e-4
    # counts only
    c3, v3 = L1.accumulate_windows(d, d_edges, H, W, None, algo="tiles")
    assert v3 is None and torch.equal(c3, counts)


def test_tiles_edge_cases(cuda_lib):
    H, W, T, B = 60, 90, 9, 5
    rec, edges = synthetic_stream(5, T, 8000, H, W, dur_ns=1_000_000)
    edges = edges.copy()
    edges[5] = edges[4]                         # an empty window
    edges[0] += 300_000                         # records before the first window
    edges[-1] -= 250_000                        # and after the last one
    # a few records that every mask drops: out-of-range coordinates and skip polarity
    bad = make_records([W, 3, 65535], [2, H, 1], [edges[2] + 5, edges[2] + 6, edges[2] + 7], [1, 0, 1])
    bad2 = make_records([4], [4], [edges[3] + 9], [2])
    rec = np.concatenate([rec, bad, bad2])
    from evfly_b200.events import records_time_ns
    rec = rec[np.argsort(records_time_ns(rec), kind="stable")]
    c_ref, v_ref = O.windows(rec, edges, H, W, B=B)
    d, d_edges = to_device(rec), torch.from_numpy(edges).cuda()
    counts, vox = L1.accumulate_windows(d, d_edges, H, W, B, algo="tiles")
    assert np.array_equal(counts.cpu().numpy(), c_ref)
    _check_voxel(vox.cpu().numpy(), v_ref)
    # no events at all
    empty = torch.empty((0, 16), dtype=torch.uint8, device="cuda")
    c0, v0 = L1.accumulate_windows(empty, d_edges, H, W, B, algo="tiles")
    assert not c0.any().item() and not v0.any().item()


def test_wire_format_device_packer_equals_host_packer_and_16_byte_path(cuda_lib):
    H, W, T, B = 260, 346, 6, 5
    rec, edges = synthetic_stream(3, T, 60_000, H, W)
    edges = edges.copy()
    edges[0] += 1_000_000
    r8_host, offs_host = pack_ev8_host(rec, edges)
    d, d_edges = to_device(rec), torch.from_numpy(edges).cuda()
    r8_dev, offs_dev = L1.pack_ev8(d, d_edges)
    assert np.array_equal(offs_dev.cpu().numpy(), offs_host)
    assert np.array_equal(r8_dev.cpu().numpy().view(EVENT8_DTYPE).reshape(-1), r8_host)          # byte for byte
    c16, v16 = L1.accumulate_windows(d, d_edges, H, W, B, algo="tiles")
    c8, v8 = L1.accumulate_windows_ev8(r8_dev, offs_dev, d_edges[:-1].contiguous(), d_edges[1:].contiguous(), H, W, B)
    assert torch.equal(c8, c16)                                                                  # counts bit-exact
    c_ref, v_ref = O.windows(rec, edges, H, W, B=B)
    assert np.array_equal(c8.cpu().numpy(), c_ref)
    _check_voxel(v8.cpu().numpy(), v_ref)


def test_wire_batch_time_major_slots(cuda_lib):
    """n_traj trajectories end to end in one buffer; window (s,t) lands in frame slot t*n_traj + s."""
    H, W, T, B, n = 260, 346, 5, 5, 3
    streams = [synthetic_stream(40 + s, T, 20_000 + 7000 * s, H, W) for s in range(n)]
    wb = WireBatch.from_streams(streams, "cuda", pin=False)
    dev = wb.on_device(wb.records.cuda())
    counts, vox = L1.accumulate_windows_ev8(dev.records, dev.win_offsets, dev.win_t0, dev.win_t1, H, W, B, out_slot=dev.out_slot, n_slots=n * T)
    for s, (rec, edges) in enumerate(streams):
        c_ref, v_ref = O.windows(rec, edges, H, W, B=B)
        assert np.array_equal(counts.view(T, n, 2, H, W)[:, s].cpu().numpy(), c_ref)
        _check_voxel(vox.view(T, n, B, H, W)[:, s].cpu().numpy(), v_ref)


@pytest.mark.parametrize("H,W,T,n_per,dur_us", [(260, 346, 6, 100_000, 33_333), (480, 640, 3, 40_000, 33_333), (260, 346, 5, 6, 33_333),
                                                 (37, 53, 4, 9_000, 2_000), (260, 346, 2, 300, 2_000_000)])
def test_4_byte_wire_format_equals_8_byte(cuda_lib, H, W, T, n_per, dur_us):
    """evfly_event4 (delta-coded microsecond time, per-chunk base table): the same frames as evfly_event8 on the same stream --
    counts bit for bit, voxel within the usual sum-order bound -- and as the oracle. 6 events per 33 ms window and 300 per 2 s
    window force escape records (gaps above 4095 us); 100 k events span 13 chunks per window."""
    rec, edges = synthetic_stream(5, T, n_per, H, W, dur_ns=dur_us * 1000, grid_ns=1000)
    edges = edges.copy()
    edges[0] += 7_000          # events before the first edge are dropped
    packed = pack_ev4_host(rec, edges)
    assert packed is not None
    r4, offs4, cb = packed
    if n_per <= 300:
        assert ((r4 & 0x7FFFF) == (1023 | (511 << 10))).any()          # escape records are present
    assert cb.shape[0] == sum(-(-int(n) // 8192) for n in np.diff(offs4))
    d4 = torch.from_numpy(r4.view(np.uint8).reshape(-1, 4)).cuda()
    t0, t1 = torch.from_numpy(edges[:-1].copy()).cuda(), torch.from_numpy(edges[1:].copy()).cuda()
    B = 5
    c4, v4 = L1.accumulate_windows_ev4(d4, torch.from_numpy(offs4).cuda(), t0, t1, torch.from_numpy(cb.view(np.int32)).cuda(), H, W, B)
    r8, offs8 = pack_ev8_host(rec, edges)
    d8 = torch.from_numpy(r8.view(np.uint8).reshape(-1, 8)).cuda()
    c8, v8 = L1.accumulate_windows_ev8(d8, torch.from_numpy(offs8).cuda(), t0, t1, H, W, B)
    assert torch.equal(c4, c8)
    assert (v4 - v8).abs().max().item() <= 1e-4
    c_ref, v_ref = O.windows(rec, edges, H, W, B=B)
    assert np.array_equal(c4.cpu().numpy(), c_ref)
    _check_voxel(v4.cpu().numpy(), v_ref)


def test_wire_batch_picks_the_4_byte_format_when_exact(cuda_lib):
    H, W, T, B, n = 260, 346, 4, 5, 3
    on_grid = [synthetic_stream(60 + s, T, 30_000, H, W, dur_ns=33_333_000, grid_ns=1000) for s in range(n)]
    wb = WireBatch.from_streams(on_grid, "cuda", pin=False)
    assert wb.record_bytes == 4 and wb.chunk_base is not None
    counts, vox = L1.accumulate_windows_wire(wb.on_device(wb.records.cuda()), H, W, B)
    for s, (rec, edges) in enumerate(on_grid):
        c_ref, v_ref = O.windows(rec, edges, H, W, B=B)
        assert np.array_equal(counts.view(T, n, 2, H, W)[:, s].cpu().numpy(), c_ref)
        _check_voxel(vox.view(T, n, B, H, W)[:, s].cpu().numpy(), v_ref)
    off_grid = [synthetic_stream(60 + s, T, 30_000, H, W) for s in range(n)]        # nanosecond timestamps
    wb8 = WireBatch.from_streams(off_grid, "cuda", pin=False)
    assert wb8.record_bytes == 8 and wb8.chunk_base is None
    with pytest.raises(ValueError):
        WireBatch.from_streams(off_grid, "cuda", pin=False, fmt=4)


def test_tiles_10M_events_one_window(cuda_lib):
    """BASELINE config 2a through the tile path: 1221 chunks in one window (two batches of the per-band chunk walk)."""
    H, W, B, n = 480, 640, 5, 10_000_000
    rec = synthetic_window(0, n, H, W)
    d = to_device(rec)
    edges = torch.tensor([0, 33_333_333], dtype=torch.int64, device="cuda")
    counts, vox = L1.accumulate_windows(d, edges, H, W, B, algo="tiles")
    c_ref, v_ref, vabs = O.voxel_window(rec, H, W, B, 0, 33_333_333, want_abs=True)
    assert np.array_equal(counts[0].cpu().numpy(), c_ref)
    _check_voxel(vox[0].cpu().numpy(), v_ref, vabs)
    assert int(counts.sum().item()) == n
    for dist_seed in (1,):
        rec = synthetic_window(dist_seed, 3_000_000, H, W, distribution="clustered")
        counts, vox = L1.accumulate_windows(to_device(rec), edges, H, W, B, algo="tiles")
        c_ref, v_ref, vabs = O.voxel_window(rec, H, W, B, 0, 33_333_333, want_abs=True)
        assert np.array_equal(counts[0].cpu().numpy(), c_ref)
        _check_voxel(vox[0].cpu().numpy(), v_ref, vabs)


def test_tiles_argument_errors(cuda_lib):
    edges = torch.tensor([0, 10], dtype=torch.int64, device="cuda")
    with pytest.raises(_lib.EvflyError):
        L1.accumulate_windows(torch.empty((0, 16), dtype=torch.uint8, device="cuda"), edges, 10, 70_000, None, algo="tiles")   # W too large
    with pytest.raises(_lib.EvflyError):
        L1.accumulate_windows(torch.empty((4, 16), dtype=torch.uint8, device="cuda"), edges, 10, 10, None, algo="tiles", sorted_by_time=False)
