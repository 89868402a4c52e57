"""Host-side helpers of the multi-GPU path: env sharding and the few collectives
PPO needs.  Pure torch.distributed, so the logic is testable with the gloo backend on CPU."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_envs(global_envs: int, world: int, rank: int):
    """GPU g owns envs [g*N/G, (g+1)*N/G); N/G must keep the env % 8 brick-type pattern (GS:962-965,1509)."""
    if global_envs % world:
        raise ValueError("num_envs must divide evenly over the ranks")
    per = global_envs // world
    if per % 8:
        raise ValueError("envs per rank must be a multiple of 8 (brick-type assignment is env_id % 8)")
    return rank * per, per


def allreduce_mean_(t: torch.Tensor, group=None):
    """in-place average over ranks (gradient / KL all-reduce)"""
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(world)
    return t


def global_moments(x: torch.Tensor, group=None):
    """(mean, unbiased std, count) of x over ALL ranks from one all-reduce of (sum, sum of squares, n)"""
    s = torch.stack([x.double().sum(), (x.double() ** 2).sum(), torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device)])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, group=group)
    n = s[2]
    mean = s[0] / n
    var = (s[1] - n * mean * mean) / (n - 1)
    return mean, var.clamp_min(0).sqrt(), n
