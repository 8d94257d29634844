"""Row N3 (rectification) on the GPU: evfly_remap_bicubic_f32 / evfly_remap_events_f32 through the
rectify_bag.py-shaped surface, bit-exact against the oracle (itself pinned on cv2.remap) and against the golden
vectors produced by the reference's Aligner on its shipped calibration."""
import os

import numpy as np
import pytest
import torch

from evfly_b200 import calibration_tools as CT
from oracle import ev_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "remap_golden.npz"))


def _evframe(u8):
    f = u8.astype(np.float32)
    f -= 128
    f *= 0.2
    return f


def test_remap_matches_reference_aligner_golden(cuda_lib, G):
    ev = _evframe(G["u8"])
    src = torch.from_numpy(ev).cuda()[None]
    u8 = torch.from_numpy(G["u8"]).cuda()[None]
    for k in ("tl", "mid", "br"):
        mx, my = torch.from_numpy(G[f"{k}_mapx"]).cuda(), torch.from_numpy(G[f"{k}_mapy"]).cuda()
        got = CT.remap_bicubic(src, mx, my)[0].cpu().numpy()
        assert np.array_equal(got, G[f"{k}_out"]), k                     # bit-exact float32
        got8 = CT.remap_bicubic(u8, mx, my)[0].cpu().numpy()              # byte image decoded on the fly
        assert np.array_equal(got8, G[f"{k}_out"]), k


@pytest.mark.parametrize("H,W,OH,OW,N", [(40, 56, 37, 45, 1), (260, 346, 260, 346, 3), (5, 7, 9, 11, 2), (3, 3, 8, 8, 1), (480, 640, 100, 90, 2)])
def test_remap_bicubic_bit_exact_vs_oracle(cuda_lib, H, W, OH, OW, N):
    rng = np.random.default_rng(H * 1000 + OW)
    src = rng.normal(0, 3, (N, H, W)).astype(np.float32)
    gx, gy = np.meshgrid(np.arange(OW, dtype=np.float32), np.arange(OH, dtype=np.float32))
    mx = (gx * (W / OW) + rng.normal(0, 2.0, (OH, OW))).astype(np.float32)     # leaves the image at the borders
    my = (gy * (H / OH) + rng.normal(0, 2.0, (OH, OW))).astype(np.float32)
    mx[0, 0], my[0, 0] = -100.0, -100.0                                       # window entirely outside
    mx[-1, -1] = 7.015625                                                     # exactly on a 1/64 tie of the 1/32 grid
    got = CT.remap_bicubic(torch.from_numpy(src).cuda(), torch.from_numpy(mx).cuda(), torch.from_numpy(my).cuda()).cpu().numpy()
    for n in range(N):
        assert np.array_equal(got[n], O.remap_bicubic(src[n], mx, my))
    for flip, rot in ((True, False), (False, True), (True, True)):
        g2 = CT.remap_img(src[0], (mx, my), flip, rot)
        assert np.array_equal(g2, O.remap_img(src[0], (mx, my), flip, rot)), (flip, rot)


def test_remap_window_of_maps_equals_crop_of_full_remap(cuda_lib):
    """run.py:339-351 aligns and then crops; the pipeline remaps only the cropped window of the maps."""
    rng = np.random.default_rng(5)
    H, W, h, w = 96, 128, 52, 70
    src = torch.from_numpy(rng.normal(0, 1, (2, H, W)).astype(np.float32)).cuda()
    gx, gy = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    mx = torch.from_numpy((gx * 0.97 + 1.3 + 0.01 * gy).astype(np.float32)).cuda()
    my = torch.from_numpy((gy * 1.02 - 0.8).astype(np.float32)).cuda()
    al = CT.Aligner(davis_map=(mx, my))
    full = CT.remap_bicubic(src, mx, my)
    wx, wy = al.davis_window(h, w)
    part = CT.remap_bicubic(src, wx, wy)
    r0, c0 = H // 2 - h // 2, W // 2 - w // 2
    assert torch.equal(part, full[:, r0:r0 + h, c0:c0 + w])
    out = al.align(davis=src[0].cpu().numpy())
    assert out["depth"] is None and np.array_equal(out["davis"], full[0].cpu().numpy())


def test_remap_events_golden(cuda_lib, G):
    H, W = (int(v) for v in G["shape"])
    # the golden events live in the top-left 64x64 window of the inverse maps; pad the window to full size
    mx, my = np.zeros((H, W), np.float32), np.zeros((H, W), np.float32)
    mx[:64, :64], my[:64, :64] = G["inv_mapx_win"], G["inv_mapy_win"]
    ev = {k: G[f"ev_in_{k}"] for k in ("x", "y", "t", "p")}
    for rot in (0, 1):
        got = CT.remap_events(ev, (mx, my), bool(rot), (W, H))
        for k in ("x", "y", "t", "p"):
            assert np.array_equal(got[k], G[f"ev_rot{rot}_{k}"]), (rot, k)
        want = O.remap_events(ev, (mx, my), bool(rot), (W, H))
        assert all(np.array_equal(got[k], want[k]) for k in want)


def test_pipeline_with_aligner_matches_oracle_chain(cuda_lib):
    """decode -> rectify -> centre crop -> quantile/clip (run.py:334-351, 250-253) vs the oracle chain."""
    from evfly_b200.events import to_device
    from evfly_b200.pipeline import PerceptionPipeline, build_deployed_model
    from evfly_b200.synthetic import synthetic_stream
    H, W, T = 480, 640, 3
    rec, edges = synthetic_stream(3, T, 60_000, H, W)
    gx, gy = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    mx = (gx + 3.0 * np.sin(gy / 50.0)).astype(np.float32)
    my = (gy * 0.99 + 2.0).astype(np.float32)
    al = CT.Aligner(davis_map=(mx, my))
    pipe = PerceptionPipeline(build_deployed_model("cuda"), sensor_hw=(H, W), aligner=al)
    frames, counts, _ = pipe.frames_from_windows(to_device(rec), torch.from_numpy(edges).cuda(), want_voxel=False)
    c = counts.cpu().numpy()
    ev = (0.2 * (c[:, 1] - c[:, 0]).astype(np.float32)).astype(np.float32)
    want = np.stack([O.remap_bicubic(ev[t], mx, my)[480 // 2 - 130:480 // 2 + 130, 640 // 2 - 173:640 // 2 + 173] for t in range(T)])
    want, _ = O.quantile_scale_clip(want[:, None], 0.97, -1.0, 1.0)
    assert np.array_equal(frames.cpu().numpy(), want)
