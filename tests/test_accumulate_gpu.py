"""GPU parity of the L1 kernels (through the C ABI) against the oracle and the golden vectors
of the reference's form_eventframe. Integer outputs are compared bit-exactly; the fp32 voxel
grid within 1e-6 relative to the per-cell sum of |weights| (north_star: 1e-6 relative)."""
import os

import numpy as np
import pytest
import torch

from evfly_b200 import _lib
from evfly_b200.events import L1, make_records, records_time_ns, to_device
from evfly_b200.synthetic import records_to_rows, synthetic_stream, synthetic_window
from oracle import ev_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "events_golden.npz"))


def _frame(view_events, H, W, **kw):
    from evfly_b200.ev_utils import form_eventframe
    return form_eventframe(view_events, H, W, **kw)


# ---- form_eventframe drop-in vs the reference's golden vectors ------------------------------
def test_form_eventframe_golden_all_events(cuda_lib, G):
    for k in ("A", "F"):
        fr = _frame(G[f"{k}_rows"], 24, 32, all_events=True)
        assert fr.dtype == np.float64 and fr.shape == (24, 32)
        assert np.array_equal(fr, G[f"{k}_frame"])          # bit-exact float64


def test_form_eventframe_golden_edges(cuda_lib, G):
    assert np.array_equal(_frame(G["B_rows"], 24, 32, all_events=True), G["B_frame_all"])
    fr, t1 = _frame(G["B_rows"], 24, 32, times0=0.0, times1=[0.02])
    assert np.array_equal(fr, G["B_frame_timed"]) and t1 == [0.02]


def test_form_eventframe_golden_timed_thresholds_and_first_n(cuda_lib, G):
    t0, t1, pt, nt = G["C_args"]
    fr, _ = _frame(G["C_rows"], 24, 32, times0=t0, times1=[t1], pos_thresh=pt, neg_thresh=nt)
    assert np.array_equal(fr, G["C_frame"])
    fr, t1o = _frame(G["C_rows"], 24, 32, times0=G["D_args"][0], N=int(G["D_args"][1]))
    assert np.array_equal(fr, G["D_frame"]) and t1o == G["D_times1"]
    fr, t1o = _frame(G["C_rows"], 24, 32, times0=0.03, N=100000)
    assert np.array_equal(fr, G["D2_frame"]) and t1o == G["D2_times1"]


def test_form_eventframe_empty_and_errors(cuda_lib, G):
    assert np.array_equal(_frame(np.zeros((0, 4)), 24, 32, all_events=True), G["E_all"])
    fr, t = _frame(np.zeros((0, 4)), 24, 32, times0=0.5, times1=[1.0])
    assert np.array_equal(fr, G["E_timed"]) and t == 0.5
    with pytest.raises(SystemExit):
        _frame(G["A_rows"], 24, 32)
    with pytest.raises(ValueError):
        _frame(G["A_rows"], 24, 32, times0=0.0)
    with pytest.raises(IndexError):
        _frame(G["A_rows"], 24, 32, times0=1e6, N=5)


def test_form_eventframe_config1_and_ragged_sizes(cuda_lib):
    for seed, n, H, W in [(0, 100_000, 260, 346), (1, 1, 260, 346), (2, 257, 3, 5), (3, 4099, 480, 640)]:
        rows = records_to_rows(synthetic_window(seed, n, H, W))
        assert np.array_equal(_frame(rows, H, W, all_events=True), O.form_eventframe(rows, H, W, all_events=True))


# ---- raw records: counts, node frames -----------------------------------------------------
def test_counts_bit_exact_uniform_and_clustered(cuda_lib):
    H, W = 480, 640
    for dist in ("uniform", "clustered"):
        rec = synthetic_window(9, 300_000, H, W, distribution=dist)
        rec["x"][:500] = 700                 # outside: node.cpp:31 bounds check
        rec["y"][500:900] = 480
        rec["polarity"][900:1000] = 2
        c = L1.accumulate_counts(to_device(rec), H, W).cpu().numpy()
        assert np.array_equal(c, O.event_counts(rec, H, W))


def test_counts_accumulate_across_calls(cuda_lib):
    H, W = 64, 96
    rec = synthetic_window(4, 50_000, H, W)
    out = torch.zeros((2, H, W), dtype=torch.int32, device="cuda")
    for part in np.array_split(rec, 7):
        L1.accumulate_counts(to_device(part), H, W, out=out)
    assert np.array_equal(out.cpu().numpy(), O.event_counts(rec, H, W))


@pytest.mark.parametrize("saturate", [False, True])
def test_image_publisher_matches_node_cpp(cuda_lib, saturate):
    from evfly_b200.accumulator import ImagePublisher
    pub = ImagePublisher(saturate=saturate)
    rng = np.random.default_rng(3)
    for window in range(3):
        rec = synthetic_window(100 + window, 120_000, 480, 640, distribution="clustered")
        # hot pixels that overflow a uint8: > 127 same-polarity events, in different orders
        hot = make_records(np.r_[np.full(400, 10), np.full(400, 11), np.full(500, 12)],
                           np.r_[np.full(400, 20), np.full(400, 20), np.full(500, 20)],
                           np.arange(1300) * 1000,
                           np.r_[np.ones(300), np.zeros(100), np.tile([1, 1, 0, 1], 100), np.zeros(500)].astype(int))
        rec = np.concatenate([rec, hot])
        rec = rec[np.argsort(records_time_ns(rec), kind="stable")]
        for part in np.array_split(rec, int(rng.integers(1, 6))):   # several EventArray messages
            pub.eventArrayCallback(part)
        img = pub.timerCallback()
        want = O.node_accumulate(rec, saturate=saturate)
        assert img.dtype == np.uint8 and img.shape == (307200,)
        assert np.array_equal(img, want)
    assert np.array_equal(pub.timerCallback(), np.full(307200, 128, np.uint8))   # reset + empty window


def test_saturate_replay_many_flagged_pixels(cuda_lib):
    # every pixel of a small sensor overflows: the flagged list has to be regrown
    from evfly_b200.accumulator import ImagePublisher
    H, W = 8, 16
    n_per = 300
    xs = np.tile(np.arange(W), H * n_per)
    ys = np.repeat(np.arange(H), W * n_per) % H
    rng = np.random.default_rng(0)
    pol = (rng.random(xs.shape[0]) < 0.8).astype(int)
    rec = make_records(xs, ys, np.arange(xs.shape[0]), pol)
    rec = rec[rng.permutation(rec.shape[0])]
    pub = ImagePublisher(saturate=True, width=W, height=H)
    pub.eventArrayCallback(rec)
    assert np.array_equal(pub.timerCallback(), O.node_accumulate(rec, True, W=W, H=H))


# ---- voxel grids ----------------------------------------------------------------------------
def _check_voxel(vox_gpu, vox_ref, vabs):
    err = np.abs(vox_gpu.astype(np.float64) - vox_ref)
    tol = 1e-6 * np.maximum(1.0, vabs)     # 1e-6 relative to the per-cell sum of |weights|
    assert (err <= tol).all(), float((err / np.maximum(1.0, vabs)).max())


@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("dist", ["uniform", "clustered"])
def test_voxel_window_vs_oracle(cuda_lib, algo, dist):
    H, W, B = 480, 640, 5
    t0, dur = 1_000_000_123, 33_333_333
    rec = synthetic_window(21, 400_000, H, W, t0_ns=t0 - 1000, dur_ns=dur + 2000, distribution=dist)
    counts, vox = L1.voxelize_window(to_device(rec), H, W, B, t0, t0 + dur, algo=algo)
    c_ref, v_ref, vabs = O.voxel_window(rec, H, W, B, t0, t0 + dur, want_abs=True)
    assert np.array_equal(counts.cpu().numpy(), c_ref)          # bit-exact
    _check_voxel(vox.cpu().numpy(), v_ref, vabs)
    # invariant tying the voxel grid to the reference frame: sum_b V[b] = npos - nneg
    s = vox.cpu().numpy().astype(np.float64).sum(0)
    assert np.abs(s - (c_ref[1].astype(float) - c_ref[0])).max() <= 1e-4


@pytest.mark.parametrize("B", [1, 2, 9])
def test_voxel_bins_and_workspace_left_zero(cuda_lib, B):
    H, W = 37, 53
    rec = synthetic_window(2, 20_000, H, W, dur_ns=10_000)
    d = to_device(rec)
    ws = L1.voxel_workspace(H, W, B, d.device)
    counts, vox = L1.voxelize_window(d, H, W, B, 0, 10_000, ws=ws, algo=1)
    c_ref, v_ref, vabs = O.voxel_window(rec, H, W, B, 0, 10_000, want_abs=True)
    assert np.array_equal(counts.cpu().numpy(), c_ref)
    _check_voxel(vox.cpu().numpy(), v_ref, vabs)
    assert not ws.any().item()                # stage is left zero for the next window


def test_windows_sorted_and_unsorted(cuda_lib):
    H, W, T, B = 120, 160, 23, 5
    rec, edges = synthetic_stream(8, T, 30_000, H, W, dur_ns=2_000_000)
    edges = edges.copy()
    edges[5] = edges[4]                        # an empty window
    c_ref, v_ref = O.windows(rec, edges, H, W, B=B)
    d_edges = torch.from_numpy(edges).cuda()
    for sorted_flag, r in ((True, rec), (False, rec[np.random.default_rng(0).permutation(rec.shape[0])])):
        counts, vox = L1.accumulate_windows(to_device(r), d_edges, H, W, B, sorted_by_time=sorted_flag)
        assert np.array_equal(counts.cpu().numpy(), c_ref)
        assert np.abs(vox.cpu().numpy() - v_ref).max() <= 1e-5
    counts, vox = L1.accumulate_windows(to_device(rec), d_edges, H, W, None)
    assert vox is None and np.array_equal(counts.cpu().numpy(), c_ref)


def test_sliced_frames_drop_in(cuda_lib):
    from evfly_b200.ev_utils import form_eventframes_sliced
    H, W, T = 60, 90, 9
    rec, edges = synthetic_stream(5, T, 8000, H, W, dur_ns=1_000_000)
    t = records_time_ns(rec)
    ev = {"x": torch.from_numpy(rec["x"].astype(np.int64)), "y": torch.from_numpy(rec["y"].astype(np.int64)),
          "t": torch.from_numpy(t), "p": torch.from_numpy(np.where(rec["polarity"] == 1, 1, -1))}
    t_edges = edges.astype(np.float64) + 0.5     # float edges as to_events.py:402-403 computes them
    got = form_eventframes_sliced(ev, t_edges, H, W)
    want = O.sliced_frames(rec["x"], rec["y"], t, ev["p"].numpy(), t_edges, H, W)
    assert got.dtype == np.float64 and np.array_equal(got, want)


# ---- full-size properties (BASELINE config 2) ------------------------------------------------
def test_full_size_10M_events_properties(cuda_lib):
    H, W, B, n = 480, 640, 5, 10_000_000
    rec = synthetic_window(0, n, H, W)
    d = to_device(rec)
    counts, vox = L1.voxelize_window(d, H, W, B, 0, 33_333_333, algo=1)
    counts2 = L1.accumulate_counts(d, H, W)
    assert torch.equal(counts, counts2)                                    # two kernels agree
    assert int(counts.sum().item()) == n                                   # every event counted once
    c = counts.cpu().numpy()
    assert np.array_equal(c, O.event_counts(rec, H, W))                    # C oracle: 10M in < 1 s
    s = vox.double().sum(0).cpu().numpy()
    assert np.abs(s - (c[1].astype(float) - c[0])).max() <= 1e-4          # sum_b V = npos - nneg
    # linearity: the two halves of the stream add up to the whole
    ca = L1.accumulate_counts(d[: n // 2], H, W)
    L1.accumulate_counts(d[n // 2:], H, W, out=ca)
    assert torch.equal(ca, counts)


# ---- L2 ----------------------------------------------------------------------------------
def test_decode_crop_and_quantile(cuda_lib):
    lib = cuda_lib
    rng = np.random.default_rng(1)
    u8 = rng.integers(100, 160, (4, 480, 640)).astype(np.uint8)
    u8[3] = 128
    u8[3, 200:260, 300:330] = 131      # sparse frame: < 3 % active -> quantile 0 (reference F8b)
    d_u8 = torch.from_numpy(u8).cuda()
    out = torch.empty((4, 1, 260, 346), dtype=torch.float32, device="cuda")
    _lib.check(lib.evfly_decode_crop(d_u8.data_ptr(), None, 4, 480, 640, 260, 346, 0.2, out.data_ptr(), _lib.stream_ptr()))
    want = O.decode_crop(u8)
    assert np.array_equal(out.cpu().numpy()[:, 0], want)
    q = torch.empty(4, dtype=torch.float32, device="cuda")
    norm = torch.empty_like(out)
    _lib.check(lib.evfly_quantile_scale_clip(out.data_ptr(), 4, 260 * 346, 0.97, -1.0, 1.0, 1e-3, norm.data_ptr(), q.data_ptr(), _lib.stream_ptr()))
    w_out, w_q = O.quantile_scale_clip(want[:, None], 0.97, -1.0, 1.0, 1e-3)
    assert np.array_equal(q.cpu().numpy(), w_q)
    assert np.array_equal(norm.cpu().numpy(), w_out, equal_nan=True)
    assert w_q[3] == 0 and np.isnan(w_out[3]).any()
    # torch.quantile on the device agrees as well (third-party arithmetic the reference calls)
    tq = torch.quantile(out.abs().flatten(1), .97, dim=1)
    assert torch.equal(tq, q)


def test_quantile_no_ties_and_counts_input(cuda_lib):
    lib = cuda_lib
    x = torch.randn((3, 1000), device="cuda")
    q = torch.empty(3, device="cuda")
    y = torch.empty_like(x)
    _lib.check(lib.evfly_quantile_scale_clip(x.data_ptr(), 3, 1000, 0.97, -1.0, 1.0, 0.0, y.data_ptr(), q.data_ptr(), _lib.stream_ptr()))
    assert torch.equal(q, torch.quantile(x.abs(), .97, dim=1))
    assert torch.equal(y, torch.clip(x / q[:, None], -1, 1))
    # counts input of decode_crop: 0.2f * (npos - nneg)
    H, W = 480, 640
    rec = synthetic_window(3, 100_000, H, W)
    counts = L1.accumulate_counts(to_device(rec), H, W)
    out = torch.empty((1, 1, 260, 346), dtype=torch.float32, device="cuda")
    _lib.check(lib.evfly_decode_crop(None, counts.data_ptr(), 1, H, W, 260, 346, 0.2, out.data_ptr(), _lib.stream_ptr()))
    img = O.node_accumulate(rec, saturate=False).reshape(H, W)
    assert np.array_equal(out.cpu().numpy()[0, 0], O.decode_crop(img))


# ---- tile-binned count accumulation == one-RED-per-event accumulation, bit for bit ---------------------
@pytest.mark.parametrize("H,W,n,dist", [(480, 640, 3_000_000, "uniform"), (480, 640, 1_500_000, "clustered"),
                                        (260, 346, 700_001, "uniform"), (33, 77, 50_000, "uniform"), (480, 640, 17, "uniform")])
def test_accumulate_counts_binned_matches_red(cuda_lib, H, W, n, dist):
    rec = synthetic_window(7, n, H, W, distribution=dist)
    rec["polarity"][::97] = 3                      # dropped records
    rec["x"][5::1001] = W + 3                 # out of frame
    d = to_device(rec)
    a = L1.accumulate_counts(d, H, W, algo="red")
    b = L1.accumulate_counts(d, H, W, algo="binned")
    assert torch.equal(a, b)
    # accumulates into an existing frame, and the workspace is reusable call after call
    b2 = L1.accumulate_counts(d, H, W, out=b.clone(), algo="binned")
    assert torch.equal(b2, 2 * a)
    assert np.array_equal(b.cpu().numpy(), O.event_counts(rec, H, W))


def test_accumulate_counts_binned_bucket_overflow_is_exact(cuda_lib):
    """Every event on one tile (here: one pixel row) overflows that tile's bucket; the overflow goes through the
    L2 reduction fallback and the result stays exact."""
    H, W, n = 480, 640, 1_200_000
    rec = synthetic_window(3, n, H, W)
    rec["y"][:] = 200
    rec["x"][: n // 2] = 17
    d = to_device(rec)
    a = L1.accumulate_counts(d, H, W, algo="red")
    b = L1.accumulate_counts(d, H, W, algo="binned")
    assert torch.equal(a, b) and int(a.sum()) == n


# ---- fused counts -> normalised frame == decode_crop + quantile_scale_clip, bit for bit -----------------
@pytest.mark.parametrize("case", ["sparse", "dense", "hot", "empty", "tiny"])
def test_counts_normalise_equals_two_step_path(cuda_lib, case):
    rng = np.random.default_rng(abs(hash(case)) % 1000)
    N, H, W, h, w = 5, 480, 640, 260, 346
    if case == "tiny":
        N, H, W, h, w = 3, 12, 16, 8, 10
    lam = {"sparse": 0.05, "dense": 3.0, "hot": 0.3, "empty": 0.0, "tiny": 1.0}[case]
    counts = rng.poisson(lam, (N, 2, H, W)).astype(np.int32)
    if case == "hot":       # more than 3 % of the pixels above 2047 events: the quantile lies in the overflow bin -> second pass
        m = rng.random((N, H, W)) < 0.08
        counts[:, 1][m] += rng.integers(2047, 9000, int(m.sum())).astype(np.int32)
    if case == "empty":
        counts[0, 1, H // 2, W // 2] = 3      # one frame with a single event, the others all zero (quantile 0 -> NaN rule)
    d = torch.from_numpy(counts).cuda()
    ref = torch.empty((N, 1, h, w), dtype=torch.float32, device="cuda")
    qref = torch.empty((N,), dtype=torch.float32, device="cuda")
    st = _lib.stream_ptr()
    _lib.check(cuda_lib.evfly_decode_crop(None, _lib.ptr(d), N, H, W, h, w, 0.2, _lib.ptr(ref), st))
    _lib.check(cuda_lib.evfly_quantile_scale_clip(_lib.ptr(ref), N, h * w, 0.97, -1.0, 1.0, 0.0, _lib.ptr(ref), _lib.ptr(qref), st))
    got = torch.empty_like(ref)
    qgot = torch.empty_like(qref)
    _lib.check(cuda_lib.evfly_counts_normalise(_lib.ptr(d), N, H, W, h, w, 0.2, 0.97, -1.0, 1.0, 0.0, _lib.ptr(got), _lib.ptr(qgot), st))
    assert np.array_equal(qgot.cpu().numpy(), qref.cpu().numpy())
    assert np.array_equal(got.cpu().numpy(), ref.cpu().numpy(), equal_nan=True)
    # and against the oracle chain
    r0, c0 = H // 2 - h // 2, W // 2 - w // 2
    fr = (0.2 * (counts[:, 1].astype(np.float32) - counts[:, 0].astype(np.float32))).astype(np.float32)[:, None, r0:r0 + h, c0:c0 + w]
    want, _ = O.quantile_scale_clip(np.ascontiguousarray(fr), 0.97, -1.0, 1.0)
    assert np.array_equal(got.cpu().numpy(), want, equal_nan=True)
