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
        feeder = TrajectoryFeeder(pipe, max_events=4 * 50_000, max_windows=4)
        got = [v.clone() for v in feeder.run(batches)]
        assert len(got) == 4
        for g, w in zip(got, want):
            assert torch.equal(g, w)          # same kernels, same inputs: bit-identical
