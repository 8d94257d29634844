"""GPU: the "next" rows N1 (dataset normalisation, learner/dataloading.py:508-533) and N2 (difflog
events, run_competition.py:603-635 / to_events.py:417-439) against the oracle."""
import numpy as np
import pytest

from evfly_b200.dataloading import normalize_event_frames
from evfly_b200.difflog import compute_events
from oracle import ev_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rescale,cutoff", [(-1.0, None), (-1.0, 1e-3), (2.5, 0.1), (0.0, 0.3)])
def test_normalize_event_frames(cuda_lib, rescale, cutoff):
    rng = np.random.default_rng(0)
    ev = (rng.integers(-20, 21, (5, 60, 90)) * 0.2 * (rng.random((5, 60, 90)) < 0.3)).astype(np.float32)
    got = normalize_event_frames(ev, rescale, cutoff).cpu().numpy()
    want = O.normalize_event_frames(ev, rescale, cutoff)
    assert np.array_equal(got, want, equal_nan=True)            # bit-exact incl. torch.quantile's interpolation


def test_difflog_events(cuda_lib):
    rng = np.random.default_rng(1)
    prev = rng.random((60, 90))
    im = np.clip(prev + rng.normal(0, 0.2, prev.shape) * (rng.random(prev.shape) < 0.5), 0, 1)
    got = compute_events(im, prev)
    want = O.difflog_events(im, prev)
    # identical algorithm in float64; CUDA's log and libm's log may differ by 1 ulp, which can only matter for a
    # difflog within 1 ulp of a multiple of the threshold
    assert got.dtype == np.float64 and np.mean(got != want) < 1e-6 and np.abs(got - want).max() <= 0.2 + 1e-12
    assert np.array_equal(got, want)
    # asymmetric thresholds, log inputs (to_events.py), below-threshold -> zeros, missing image -> zeros
    assert np.array_equal(compute_events(np.log(im + 1e-5), np.log(prev + 1e-5), 0.15, 0.3, inputs_are_log=True),
                          O.difflog_events(np.log(im + 1e-5), np.log(prev + 1e-5), 0.15, 0.3, inputs_are_log=True))
    assert not compute_events(prev * 1.01, prev).any()
    assert compute_events(None, prev, shape=(60, 90)).shape == (60, 90)


def test_resize_trajectory_matches_torch_interpolate(cuda_lib):
    """dataloading.py:401-416 (resize_input): bilinear, align_corners=False, per trajectory."""
    import torch
    import torch.nn.functional as F
    from evfly_b200.dataloading import resize_trajectory
    g = torch.Generator().manual_seed(5)
    for (T, H, W, size) in [(7, 260, 346, (60, 90)), (3, 480, 640, (260, 346)), (2, 60, 90, (60, 90)), (4, 33, 47, (64, 100))]:
        x = torch.randn(T, H, W, generator=g)
        want = F.interpolate(x.unsqueeze(1), size=size, mode="bilinear", align_corners=False).squeeze(1)
        got = resize_trajectory(x.numpy(), size).cpu()
        assert got.shape == want.shape
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)        # fp32 path tolerance (north_star)


def test_dataloading_matches_reference_dataloader_golden(cuda_lib, golden_dir):
    """normalize_event_frames / resize_trajectory against what the REFERENCE dataloader (learner/dataloading.py run on a
    synthetic folder dataset, tests/golden/make_golden_dataloading.py) returned."""
    import os
    import numpy as np
    from evfly_b200.dataloading import normalize_event_frames, resize_trajectory
    G = np.load(os.path.join(golden_dir, "dataloading_golden.npz"))
    modes = {"q97": (-1.0, None), "q97_cut": (-1.0, 0.05), "div2_cut": (2.0, 0.05), "raw": (0.0, None)}
    for k in range(2):
        for mode, (rescale, cutoff) in modes.items():
            got = normalize_event_frames(G[f"native_in{k}"], rescale, cutoff).cpu().numpy()
            assert np.array_equal(got, G[f"native_{mode}_out{k}"], equal_nan=True), (k, mode)      # bit-exact incl. the NaN frame
        got = resize_trajectory(G[f"resized_in{k}"], (60, 90)).cpu().numpy()
        np.testing.assert_allclose(got, G[f"resized_raw_out{k}"], rtol=1e-5, atol=1e-6)             # fp32 path tolerance
