"""world_size-2 gloo tests of the multi-GPU host logic (no GPU): env sharding, gradient averaging
(N-rank == 1-rank gradient equivalence), global advantage moments."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from seqdex_b200.dist_utils import global_moments, shard_envs
    from seqdex_b200.ppo import A2CAgent

    import types
    _agent = types.SimpleNamespace(dist=dist.group.WORLD, world=world)     # the two attributes A2CAgent._allreduce reads; the method is the product's
    _Agent = lambda: _agent
    allreduce_mean_ = lambda t: A2CAgent._allreduce(_agent, t, avg=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(20, 32), torch.nn.ELU(), torch.nn.Linear(32, 3))
    x, y = torch.randn(64, 20), torch.randn(64, 3)
    # full-batch gradient (what one rank with all envs computes)
    net.zero_grad()
    ((net(x) - y) ** 2).mean().backward()
    full = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
    # sharded: each rank its env slice, mean loss over the shard, then ONE all-reduce(mean) of the flat gradient
    start, per = shard_envs(64, world, rank)
    net.zero_grad()
    ((net(x[start:start + per]) - y[start:start + per]) ** 2).mean().backward()
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    allreduce_mean_(flat)
    adv = torch.randn(64) * 3 + 1
    mean, std, n = global_moments(adv[start:start + per])
    cnt = torch.ones(3)
    A2CAgent._allreduce(_Agent(), cnt, avg=False)        # the sum form (advantage / RunningMeanStd moments)
    from seqdex_b200.dist_utils import params_digest
    d = params_digest(flat)
    ds = [torch.zeros_like(d) for _ in range(world)]
    dist.all_gather(ds, d)
    ok = (torch.allclose(flat, full, atol=1e-6), bool((cnt == world).all()), all(torch.equal(ds[0], x) for x in ds), abs(float(mean) - float(adv.double().mean())) < 1e-9,
          abs(float(std) - float(adv.double().std())) < 1e-9, int(n) == 64)
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equivalence_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(all(ok) for _, ok in res), res


def test_shard_envs_rules():
    from seqdex_b200.dist_utils import shard_envs
    assert shard_envs(16384, 8, 3) == (3 * 2048, 2048)
    with pytest.raises(ValueError):
        shard_envs(100, 8, 0)
    with pytest.raises(ValueError):
        shard_envs(24, 2, 0)     # 12 per rank breaks the env % 8 pattern
