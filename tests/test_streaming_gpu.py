"""GPU: the CUDA-graph streaming session (batch-1, state carried) returns what the eager pipeline
returns, window after window, including windows of different event counts."""
import json
import os

import numpy as np
import pytest
import torch

import evfly_b200
from evfly_b200.events import to_device
from evfly_b200.pipeline import PerceptionPipeline, StreamingSession, build_deployed_model
from evfly_b200.synthetic import synthetic_window
from oracle.synth_ckpt import shapes_of, synth_state_dict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graph_replay_equals_eager(cuda_lib, precision):
    with torch.no_grad():
        m = build_deployed_model("cpu")
        m.load_state_dict(synth_state_dict(shapes_of(m), 31))
        m = evfly_b200.set_precision(m.cuda().eval(), precision)
        pipe = PerceptionPipeline(m, sensor_hw=(480, 640), model_hw=(260, 346))
        eager = PerceptionPipeline(m, sensor_hw=(480, 640), model_hw=(260, 346))
        sess = StreamingSession(pipe, capacity=131072)
        edges = torch.tensor([0, 33_333_333], dtype=torch.int64, device="cuda")
        for k, n in enumerate([100_000, 60_000, 100_000, 131_072, 90_000]):
            rec = to_device(synthetic_window(50 + k, n, 480, 640))
            vel = sess.step(rec).clone()
            evel, edep, ecounts, evox = eager(rec, edges)
            assert torch.equal(sess.counts, ecounts)                       # integer path: bit-exact
            if precision == "fp32":
                np.testing.assert_allclose(vel.cpu().numpy(), evel.cpu().numpy(), rtol=1e-5, atol=1e-6)
                np.testing.assert_allclose(sess.depth.cpu().numpy(), edep.cpu().numpy(), rtol=1e-5, atol=1e-5)
            else:
                np.testing.assert_allclose(vel.cpu().numpy(), evel.cpu().numpy(), rtol=1e-3, atol=1e-4)
        sess.reset(); eager.reset()
        rec = to_device(synthetic_window(77, 100_000, 480, 640))
        assert torch.allclose(sess.step(rec), eager(rec, edges)[0], rtol=1e-3, atol=1e-4)
