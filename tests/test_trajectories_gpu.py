"""GPU: the trajectory-batched forward (config 4 extension) gives, per trajectory, what the reference-style
single-sequence forward gives -- for the exact fp32 path (rtol 1e-5) and the bf16 path."""
import numpy as np
import pytest
import torch

import evfly_b200
from evfly_b200.events import to_device
from evfly_b200.pipeline import PerceptionPipeline, build_deployed_model
from evfly_b200.synthetic import synthetic_stream
from oracle.synth_ckpt import shapes_of, synth_state_dict, synthetic_frames

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_batched_trajectories_equal_sequential(cuda_lib, precision):
    with torch.no_grad():
        m = build_deployed_model("cpu")
        m.load_state_dict(synth_state_dict(shapes_of(m), 31))
        m = evfly_b200.set_precision(m.cuda().eval(), precision)
        n, T = 3, 5
        frames = torch.stack([synthetic_frames(40 + s, T) for s in range(n)]).cuda()          # [n,T,1,H,W]
        dv = torch.full((T, 1), 4.0, device="cuda")
        seq = [m([frames[s].clone(), dv, [None, None], None]) for s in range(n)]
        tm = frames.transpose(0, 1).reshape(T * n, 1, 260, 346).contiguous()
        vel, (dep, yu, ((hu, _), hv)) = m.forward_trajectories([tm, dv.repeat(n, 1), [None, None], None], n)
        tol = dict(rtol=1e-5, atol=2e-5) if precision == "fp32" else dict(rtol=2e-2, atol=2e-2)
        for s in range(n):
            svel, (sdep, _, ((shu, _), shv)) = seq[s]
            np.testing.assert_allclose(vel.view(T, n, 3)[:, s].cpu().numpy(), svel.cpu().numpy(), **tol)
            np.testing.assert_allclose(dep.view(T, n, 1, 260, 346)[:, s].cpu().numpy(), sdep.cpu().numpy(), **tol)
            np.testing.assert_allclose(hu[0][0][s].cpu().numpy(), shu[0][0][0].cpu().numpy(), **tol)
            np.testing.assert_allclose(hu[0][1][s].cpu().numpy(), shu[0][1][0].cpu().numpy(), **tol)
            # bf16: kernels are chosen by batch size (e.g. the patch embedding runs on the tensor cores from 1024 tokens
            # up), so batched and sequential runs agree to bf16 noise, not bit for bit; the LSTM state gets a wider atol
            np.testing.assert_allclose(hv[0][:, s].cpu().numpy(), shv[0].cpu().numpy(), **(tol if precision == "fp32" else dict(rtol=2e-2, atol=5e-2)))
        # carried state: a second chunk of the same trajectories continues where the first stopped
        frames2 = torch.stack([synthetic_frames(60 + s, T) for s in range(n)]).cuda()
        tm2 = frames2.transpose(0, 1).reshape(T * n, 1, 260, 346).contiguous()
        vel2, _ = m.forward_trajectories([tm2, dv.repeat(n, 1), [hu, None], hv], n)
        s = 1
        _, (_, _, ((shu, _), shv)) = seq[s]
        svel2, _ = m([frames2[s].clone(), dv, [shu, None], shv])
        np.testing.assert_allclose(vel2.view(T, n, 3)[:, s].cpu().numpy(), svel2.cpu().numpy(), **tol)


def test_pipeline_run_trajectories(cuda_lib):
    with torch.no_grad():
        m = build_deployed_model("cpu")
        m.load_state_dict(synth_state_dict(shapes_of(m), 31))
        m = evfly_b200.set_precision(m.cuda().eval(), "bf16")
        pipe = PerceptionPipeline(m, sensor_hw=(260, 346))
        recs, edges = [], []
        for s in range(2):
            r, e = synthetic_stream(30 + s, 4, 60_000, 260, 346)
            recs.append(to_device(r)); edges.append(torch.from_numpy(e).cuda())
        vel, depth = pipe.run_trajectories(recs, edges)
        assert vel.shape == (2, 4, 3) and depth.shape == (2, 4, 1, 260, 346)
        for s in range(2):
            pipe.reset()
            v, d, _, _ = pipe(recs[s], edges[s])
            np.testing.assert_allclose(vel[s].cpu().numpy(), v.cpu().numpy(), rtol=2e-2, atol=2e-3)


def test_prefetched_two_stage_execution_equals_run_trajectories(cuda_lib):
    """L1+L2 on the side stream (prefetch_trajectories) + model on the main stream (run_prefetched), with the next
    batch prefetched before the current model runs, gives exactly what run_trajectories gives batch by batch."""
    with torch.no_grad():
        m = build_deployed_model("cpu")
        m.load_state_dict(synth_state_dict(shapes_of(m), 31))
        m = evfly_b200.set_precision(m.cuda().eval(), "bf16")
        pipe = PerceptionPipeline(m, sensor_hw=(260, 346))
        batches = []
        for b in range(3):
            recs, edges = [], []
            for s in range(2):
                r, e = synthetic_stream(70 + 10 * b + s, 4, 50_000, 260, 346)
                recs.append(to_device(r)); edges.append(torch.from_numpy(e).cuda())
            batches.append((recs, edges))
        want = [tuple(t.clone() for t in pipe.run_trajectories(*b)) for b in batches]
        got = []
        h = pipe.prefetch_trajectories(*batches[0])
        for i in range(len(batches)):
            h_next = pipe.prefetch_trajectories(*batches[i + 1]) if i + 1 < len(batches) else None
            got.append(tuple(t.clone() for t in pipe.run_prefetched(h)))
            h = h_next
        torch.cuda.synchronize()
        for (v0, d0), (v1, d1) in zip(want, got):
            assert torch.equal(v0, v1) and torch.equal(d0, d1)
