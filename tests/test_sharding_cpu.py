"""CPU, world_size 2, gloo: the N>1 host logic (trajectory sharding + output gather) that bench.py's
multi-GPU runs and offline evaluation rely on. No compute kernels are called here."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from evfly_b200.sharding import evaluate_trajectories, shard_indices


def test_shards_partition_the_work():
    for n in (0, 1, 7, 2048):
        for world in (1, 2, 3, 8):
            seen = sorted(i for r in range(world) for i in shard_indices(n, r, world))
            assert seen == list(range(n))
            sizes = [len(shard_indices(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def _worker(rank, world, port, n_traj, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a stand-in for PerceptionPipeline: the output encodes the trajectory index
        run = lambda i: torch.full((5, 3), float(i)) + torch.arange(3.0)
        out = evaluate_trajectories(run, n_traj, rank, world)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, out.clone(), float(t.item())))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_traj", [7, 8])
def test_world_size_2_gloo_gather(n_traj):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_traj, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.stack([torch.full((5, 3), float(i)) + torch.arange(3.0) for i in range(n_traj)])
    for rank, out, tmax in results:
        assert torch.equal(out, want), f"rank {rank} gathered the wrong order"
        assert tmax == 11.0
