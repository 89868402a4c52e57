"""Host-side helpers of the multi-GPU path: env sharding (bench.py) and the collectives of ``A2CAgent``
(``ppo.py::_allreduce``).  Pure torch.distributed, so the logic is testable with the gloo backend on CPU."""
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


def allreduce_(t: torch.Tensor, group=None, avg=True):
    """in-place sum (avg: mean) over ranks -- the collective ``A2CAgent._allreduce`` issues for gradients + KL, advantage moments and
    RunningMeanStd column moments"""
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        if avg:
            t.div_(world)
    return t


def allreduce_mean_(t: torch.Tensor, group=None):
    return allreduce_(t, group, True)


def params_digest(*tensors):
    """64-bit digest of parameter tensors (sum of the fp32 bit patterns as int64, and of their squares mod 2^63): equal on every rank
    iff the replicas are in lock-step; bench.py all-gathers it after the timed region"""
    acc = torch.zeros(2, dtype=torch.int64, device=tensors[0].device)
    for t in tensors:
        b = t.detach().contiguous().view(torch.int32).to(torch.int64)
        acc[0] += b.sum()
        acc[1] += (b * (b & 0xFFFF)).sum()
    return acc


def global_moments(x: torch.Tensor, group=None):
    """(mean, unbiased std, count) of x over ALL ranks from one all-reduce of (sum, sum of squares, n)"""
    s = torch.stack([x.double().sum(), (x.double() ** 2).sum(), torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device)])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, group=group)
    n = s[2]
    mean = s[0] / n
    var = (s[1] - n * mean * mean) / (n - 1)
    return mean, var.clamp_min(0).sqrt(), n
