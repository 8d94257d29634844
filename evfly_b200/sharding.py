"""Multi-GPU offline trajectory evaluation (BASELINE config 4; learner/evaluation_tools.py:62-66
runs trajectories one after another on one device). Trajectories carry private recurrent state
and never interact, so rank r of G takes trajectories r::G with a full weight
replica and NO collective on the data path; one all_gather of the per-rank velocity commands at
the end. A single trajectory is never split across GPUs."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> list[int]:
    """Round-robin shard r::G (balanced to within one item for any n_items, world)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_items, world))


def gather_trajectory_outputs(local: torch.Tensor, n_items: int, rank: int, world: int) -> torch.Tensor:
    """local [n_local, T, D] for trajectories shard_indices(n_items, rank, world) -> [n_items, T, D] on
    every rank, in trajectory order. Uses the default process group (NCCL on GPUs, gloo in tests)."""
    if world == 1:
        return local
    n_max = (n_items + world - 1) // world
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.empty((n_items,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        out[idx] = parts[r][: len(idx)]
    return out


def evaluate_trajectories(run_trajectory, n_trajectories: int, rank: int = 0, world: int = 1) -> torch.Tensor:
    """run_trajectory(i) -> [T, D] tensor (e.g. PerceptionPipeline on trajectory i with fresh state).
    Returns the gathered [n_trajectories, T, D]."""
    mine = shard_indices(n_trajectories, rank, world)
    outs = [run_trajectory(i) for i in mine]
    if outs:
        local = torch.stack(outs)
    else:
        probe = run_trajectory(0)
        local = probe.new_zeros((0,) + tuple(probe.shape))
    return gather_trajectory_outputs(local, n_trajectories, rank, world)
