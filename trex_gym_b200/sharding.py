"""Multi-GPU sharding of environments.

Environments are independent: rank g owns global env ids ``[g*N, (g+1)*N)`` and there is no
collective on the step path.  The only exchange is an optional all-reduce of a handful of episode
statistics per logging interval (NCCL on GPUs, gloo in the CPU test-suite).
"""
from __future__ import annotations

import torch

STAT_FIELDS = ("env_steps", "episodes", "nan_resets", "sum_solver_iterations", "sum_contacts", "sum_reward",
               "sum_lifting_penalty", "sum_energy_penalty")


def shard_range(rank: int, world_size: int, envs_per_gpu: int):
    """Global env ids owned by ``rank`` (weak scaling: ``envs_per_gpu`` per rank)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    lo = rank * envs_per_gpu
    return lo, lo + envs_per_gpu


def max_over_ranks(seconds: float, device=None) -> float:
    """Device time of a timed region as the max over ranks (never wall clock of one rank)."""
    import torch.distributed as dist

    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_episode_stats(local: dict, device=None) -> dict:
    """Sum the per-rank statistics over all ranks (8 scalars; the only collective in the framework)."""
    import torch.distributed as dist

    t = torch.tensor([float(local.get(k, 0.0)) for k in STAT_FIELDS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(STAT_FIELDS, t.tolist())}
