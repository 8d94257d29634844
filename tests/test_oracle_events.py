"""CPU: the oracle (oracle/ev_oracle.py + ev_oracle.c) against the golden vectors produced by
the reference's own form_eventframe, against numpy.histogram2d / torch.quantile (the
third-party arithmetic the reference calls), and self-consistency of its C and numpy halves."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import ev_oracle as O
from evfly_b200.events import EVENT_DTYPE, make_records, records_time_ns
from evfly_b200.synthetic import records_to_rows, synthetic_stream, synthetic_window


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "events_golden.npz"))


def test_record_layout_matches_ros_struct():
    # dv_ros_msgs/msg/Event.msg:2-5 in memory: x@0 y@2 ts.sec@4 ts.nsec@8 polarity@12, 16 bytes
    assert EVENT_DTYPE.itemsize == 16 and O.EVENT_DTYPE == EVENT_DTYPE
    assert [EVENT_DTYPE.fields[k][1] for k in ("x", "y", "ts_sec", "ts_nsec", "polarity")] == [0, 2, 4, 8, 12]


def test_golden_all_events(G):
    assert np.array_equal(O.form_eventframe(G["A_rows"], 24, 32, all_events=True), G["A_frame"])
    assert np.array_equal(O.form_eventframe(G["F_rows"], 24, 32, all_events=True), G["F_frame"])


def test_golden_edge_coordinates_and_polarities(G):
    assert np.array_equal(O.form_eventframe(G["B_rows"], 24, 32, all_events=True), G["B_frame_all"])
    fr, _ = O.form_eventframe(G["B_rows"], 24, 32, times0=0.0, times1=[0.02])
    assert np.array_equal(fr, G["B_frame_timed"])


def test_golden_timed_and_thresholds(G):
    t0, t1, pt, nt = G["C_args"]
    fr, t1_out = O.form_eventframe(G["C_rows"], 24, 32, times0=t0, times1=[t1], pos_thresh=pt, neg_thresh=nt)
    assert np.array_equal(fr, G["C_frame"]) and t1_out[0] == G["C_times1"][0]


def test_golden_first_n_mode(G):
    fr, t1 = O.form_eventframe(G["C_rows"], 24, 32, times0=G["D_args"][0], N=int(G["D_args"][1]))
    assert np.array_equal(fr, G["D_frame"]) and t1 == G["D_times1"]
    fr, t1 = O.form_eventframe(G["C_rows"], 24, 32, times0=0.03, N=100000)
    assert np.array_equal(fr, G["D2_frame"]) and t1 == G["D2_times1"]


def test_golden_empty(G):
    assert np.array_equal(O.form_eventframe(np.zeros((0, 4)), 24, 32, all_events=True), G["E_all"])
    fr, t = O.form_eventframe(np.zeros((0, 4)), 24, 32, times0=0.5, times1=[1.0])
    assert np.array_equal(fr, G["E_timed"]) and t == G["E_timed_t"]


def test_golden_config1_digest(G):
    seed, n, H, W = (int(v) for v in G["G_seed_n_H_W"])
    rng = np.random.default_rng(seed)
    rows = np.empty((n, 4))
    rows[:, 0] = np.sort(rng.integers(0, 33_333_333, n))
    rows[:, 1] = rng.integers(0, W, n)
    rows[:, 2] = rng.integers(0, H, n)
    rows[:, 3] = rng.integers(0, 2, n)
    fr = O.form_eventframe(rows, H, W, all_events=True)
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(fr).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, G["G_sha256"])
    # the reference returns a transposed view, so its summation order differs: sums are approximate
    assert np.allclose([np.abs(fr).sum(), fr.sum()], G["G_sum_abs"][:2], rtol=1e-12)
    assert fr.max() == G["G_sum_abs"][2] and fr.min() == G["G_sum_abs"][3]


def test_hist2d_restatement_equals_numpy_histogram2d():
    rng = np.random.default_rng(7)
    H, W = 13, 17
    xs = rng.uniform(-2, W + 2, 5000)
    ys = rng.uniform(-2, H + 2, 5000)
    xs[:50], ys[50:100] = W, H            # right edges
    xs[100:110] = np.nan
    ref = np.histogram2d(xs, ys, bins=(W, H), range=[[0, W], [0, H]])[0].T
    assert np.array_equal(O.hist2d_counts(xs, ys, H, W), ref.astype(np.int64))


def test_c_counts_equal_numpy_restatement():
    H, W = 48, 64
    rec = synthetic_window(5, 20000, H, W)
    rec["x"][:100] = W          # out of range for the unsigned compare of node.cpp:31
    rec["polarity"][100:150] = 2
    c = O.event_counts(rec, H, W)
    keep = (rec["x"] < W) & (rec["y"] < H) & (rec["polarity"] < 2)
    rows = records_to_rows(rec[keep])
    assert np.array_equal(c, O.event_counts_rows(rows, H, W, neg_is_zero=True))
    assert c.sum() == keep.sum()


def test_node_wrap_is_mod_256_and_saturate_is_order_dependent():
    H, W = 4, 4
    n = 700
    rec = make_records(np.full(n, 1), np.full(n, 2), np.arange(n), np.ones(n, dtype=int))
    img = O.node_accumulate(rec, saturate=False, W=W, H=H)
    assert img[2 * W + 1] == (128 + n) % 256 and (np.delete(img, 2 * W + 1) == 128).all()
    img = O.node_accumulate(rec, saturate=True, W=W, H=H)
    assert img[2 * W + 1] == 255
    # 200 up then 100 down saturates at 255 -> 155; interleaved never saturates -> 228
    up_down = make_records(np.zeros(300), np.zeros(300), np.arange(300), np.r_[np.ones(200), np.zeros(100)].astype(int))
    inter = make_records(np.zeros(300), np.zeros(300), np.arange(300), np.tile([1, 1, 0], 100))
    assert O.node_accumulate(up_down, True, W, H)[0] == 155
    assert O.node_accumulate(inter, True, W, H)[0] == 228
    assert O.node_accumulate(up_down, False, W, H)[0] == O.node_accumulate(inter, False, W, H)[0] == 228


def test_voxel_sums_to_signed_count_frame():
    # the invariant that ties the build-defined voxel grid to the reference:
    # sum_b V[b] = npos - nneg = form_eventframe / 0.2
    H, W, B = 30, 40, 5
    rec = synthetic_window(11, 30000, H, W, t0_ns=5_000, dur_ns=1_000_000)
    counts, vox = O.voxel_window(rec, H, W, B, 5_000, 1_005_000)
    assert np.allclose(vox.sum(0), counts[1].astype(float) - counts[0], atol=1e-9)
    frame = O.form_eventframe(records_to_rows(rec), H, W, all_events=True)
    assert np.allclose(vox.sum(0) * 0.2, frame, atol=1e-9)
    # an event exactly at t0 lands entirely in bin 0; one at the window centre entirely in bin 2
    one = make_records([3, 4], [2, 2], [5_000, 505_000], [1, 0])
    _, v = O.voxel_window(one, H, W, B, 5_000, 1_005_000)
    assert v[0, 2, 3] == 1.0 and v[2, 2, 4] == -1.0 and np.abs(v).sum() == 2.0


def test_windows_equal_literal_to_events_slicing():
    H, W, T = 20, 28, 6
    rec, edges = synthetic_stream(3, T, 4000, H, W, dur_ns=1_000_000)
    counts, vox = O.windows(rec, edges, H, W, B=3)
    t = records_time_ns(rec)
    p = np.where(rec["polarity"] == 1, 1, -1)
    frames = O.sliced_frames(rec["x"], rec["y"], t, p, edges, H, W)
    assert np.array_equal(frames, 0.2 * counts[:, 1].astype(float) - 0.2 * counts[:, 0])
    for w in range(T):
        c1, v1 = O.voxel_window(rec, H, W, 3, edges[w], edges[w + 1])
        assert np.array_equal(c1, counts[w]) and np.array_equal(v1, vox[w])


def test_decode_crop_matches_run_py_indices():
    u8 = np.random.default_rng(0).integers(0, 256, (480, 640), dtype=np.uint8)
    f = O.decode_crop(u8)
    assert f.shape == (260, 346) and f.dtype == np.float32
    assert f[0, 0] == np.float32(np.float32(float(u8[110, 147]) - 128) * np.float32(0.2))


@pytest.mark.parametrize("seed,n", [(0, 89960), (1, 1000), (2, 37), (3, 2)])
def test_quantile_restatement_equals_torch_quantile(seed, n):
    rng = np.random.default_rng(seed)
    x = (rng.integers(-12, 13, n) * np.float32(0.2)).astype(np.float32)   # event-frame-like ties
    y = rng.normal(size=n).astype(np.float32)                             # no ties
    for v in (x, y):
        want = torch.quantile(torch.from_numpy(v).abs(), .97).numpy()
        got = O.quantile_f32(np.abs(v), 0.97)
        assert got.tobytes() == want.tobytes(), (got, want)


def test_quantile_scale_clip_matches_torch_pipeline():
    rng = np.random.default_rng(5)
    fr = (rng.integers(-9, 10, (3, 1, 26, 34)) * np.float32(0.2)).astype(np.float32)
    fr[2] = 0  # sparse frame: quantile 0 -> 0/0 NaN (reference F8b)
    out, qs = O.quantile_scale_clip(fr, cutoff=1e-3)
    for i in range(3):
        t = torch.from_numpy(fr[i:i + 1])
        s = torch.quantile(t.abs(), .97)                  # run.py:250
        ref = torch.clip(t / s, -1.0, 1.0)                # run.py:253
        ref[ref.abs() < 1e-3] = 0.0                       # learner_models.py:477
        assert np.array_equal(out[i:i + 1], ref.numpy(), equal_nan=True)
    assert np.isnan(out[2]).all()


# ---- row N3: the remap oracle vs the reference's Aligner (golden) and vs cv2 when it is importable ---------
def test_remap_oracle_matches_reference_golden(golden_dir):
    G = np.load(os.path.join(golden_dir, "remap_golden.npz"))
    ev = G["u8"].astype(np.float32)
    ev -= 128
    ev *= 0.2
    for k in ("tl", "mid", "br"):
        assert np.array_equal(O.remap_bicubic(ev, G[f"{k}_mapx"], G[f"{k}_mapy"]), G[f"{k}_out"]), k
    H, W = (int(v) for v in G["shape"])
    mx, my = np.zeros((H, W), np.float32), np.zeros((H, W), np.float32)
    mx[:64, :64], my[:64, :64] = G["inv_mapx_win"], G["inv_mapy_win"]
    evs = {k: G[f"ev_in_{k}"] for k in ("x", "y", "t", "p")}
    for rot in (0, 1):
        got = O.remap_events(evs, (mx, my), bool(rot), (W, H))
        for k in ("x", "y", "t", "p"):
            assert np.array_equal(got[k], G[f"ev_rot{rot}_{k}"]), (rot, k)


def test_remap_oracle_matches_cv2_bit_for_bit():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for (H, W, OH, OW) in [(40, 56, 37, 45), (120, 160, 120, 160), (5, 7, 9, 11), (3, 3, 8, 8)]:
        src = rng.normal(0, 3, (H, W)).astype(np.float32)
        mx = (rng.random((OH, OW)) * (W + 10) - 5).astype(np.float32)
        my = (rng.random((OH, OW)) * (H + 10) - 5).astype(np.float32)
        assert np.array_equal(O.remap_bicubic(src, mx, my), cv2.remap(src, mx, my, cv2.INTER_CUBIC))
        assert np.array_equal(O.remap_img(src, (mx, my), True, True),
                              cv2.rotate(cv2.remap(np.ascontiguousarray(src[:, ::-1]), mx, my, cv2.INTER_CUBIC), cv2.ROTATE_180))


# ---- rows N1 / N4: the oracle vs what the REFERENCE dataloader returned for a synthetic folder dataset -----
def test_normalisation_oracle_matches_reference_dataloader_golden(golden_dir):
    G = np.load(os.path.join(golden_dir, "dataloading_golden.npz"))
    modes = {"q97": (-1.0, None), "q97_cut": (-1.0, 0.05), "div2_cut": (2.0, 0.05), "raw": (0.0, None)}
    for k in range(2):
        x = G[f"native_in{k}"]
        for mode, (rescale, cutoff) in modes.items():
            got = O.normalize_event_frames(x, rescale, cutoff)
            assert np.array_equal(got, G[f"native_{mode}_out{k}"], equal_nan=True), (k, mode)
        # resize_input step (dataloading.py:401-416) followed by the normalisation
        xr = torch.nn.functional.interpolate(torch.from_numpy(G[f"resized_in{k}"]).unsqueeze(1), size=(60, 90), mode="bilinear",
                                             align_corners=False).squeeze().numpy()
        assert np.array_equal(xr, G[f"resized_raw_out{k}"])
        assert np.array_equal(O.normalize_event_frames(xr, -1.0, 0.05), G[f"resized_q97_cut_out{k}"], equal_nan=True)
    assert G["native_desvel"].tolist() == [4.0, 4.0, 4.0, 5.0, 5.0, 5.0] and G["native_lengths"].tolist() == [3, 3]
    assert np.isnan(G["native_q97_out0"][0]).all()        # the all-zero frame: quantile 0 -> 0/0 (reference behaviour F8b)
