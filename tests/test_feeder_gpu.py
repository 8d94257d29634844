"""GPU: the double-buffered end-to-end feeder returns, per trajectory, what the plain pipeline returns."""
import numpy as np
import pytest
import torch

import evfly_b200
from evfly_b200.pipeline import PerceptionPipeline, TrajectoryFeeder, build_deployed_model
from evfly_b200.synthetic import synthetic_stream
from oracle.synth_ckpt import shapes_of, synth_state_dict

pytestmark = pytest.mark.gpu


def test_feeder_matches_pipeline(cuda_lib):
    with torch.no_grad():
        m = build_deployed_model("cpu")
        m.load_state_dict(synth_state_dict(shapes_of(m), 31))
        m = evfly_b200.set_precision(m.cuda().eval(), "bf16")
        pipe = PerceptionPipeline(m, sensor_hw=(260, 346))
        batches, want = [], []
        for k, T in enumerate([3, 2, 4, 1]):
            rec, edges = synthetic_stream(20 + k, T, 50_000, 260, 346)
            pinned = torch.from_numpy(rec.view(np.uint8).reshape(-1, 16)).pin_memory()
            d_edges = torch.from_numpy(edges).cuda()
            batches.append((pinned, d_edges))
            pipe.reset()
            want.append(pipe(pinned.cuda(), d_edges)[0].cpu())
        feeder = TrajectoryFeeder(pipe, max_events=6 * 50_000, max_windows=6)
        got = [v.clone() for v in feeder.run(batches)]
        assert len(got) == 4
        for g, w in zip(got, want):
            assert torch.equal(g[0], w)          # same kernels, same inputs: bit-identical
        # multi-trajectory batches advance together
        multi = [([batches[0][0], batches[0][0]], [batches[0][1], batches[0][1]])]
        (v,) = [v.clone() for v in feeder.run(multi)]
        assert v.shape == (2, 3, 3) and torch.allclose(v[0], want[0], rtol=2e-2, atol=2e-3) and torch.equal(v[0], v[1])
