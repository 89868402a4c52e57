#!/usr/bin/env python
"""A2CAgent.update eager vs CUDA graph: wall time per update and bit-equality of the results, on the same rollouts.
Usage: tools/ppo_graph_check.py [num_envs] [minibatch]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    mb = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    from seqdex_b200.ppo import A2CAgent, PPOConfig
    from seqdex_b200.tasks import BlockAssemblyGraspSim
    from seqdex_b200.vec_task import RLgamesVecTaskPython
    cfg = {"env": {"numEnvs": n, "episodeLength": 150, "actionsMovingAverage": 1.0}, "sim": {"substeps": 2, "physx": {}}, "task": {"randomize": False}}
    res = {}
    for mode in ("0", "force"):
        os.environ["SEQDEX_PPO_GRAPH"] = mode
        torch.manual_seed(17)
        task = BlockAssemblyGraspSim(cfg, bank_per_type=2)
        agent = A2CAgent(RLgamesVecTaskPython(task, "cuda:0"), PPOConfig(minibatch_size=min(mb, 8 * n)))
        snaps, times = [], []
        for it in range(6):
            agent.play_steps()
            torch.cuda.synchronize()
            t0 = time.time()
            info = agent.update()
            torch.cuda.synchronize()
            times.append((time.time() - t0) * 1e3)
            snaps.append((agent.actor.params.clone(), agent.cv.params.clone(), info))
        res[mode] = snaps
        print(f"mode {mode}: update ms per call", [round(t, 2) for t in times], "graph" if getattr(agent, "_graph", None) is not None else "eager",
              getattr(agent, "_graph_launches", None))
        task.env.close()
    for it, (a, b) in enumerate(zip(res["0"], res["force"])):
        print(it, "actor max diff", float((a[0] - b[0]).abs().max()), "cv max diff", float((a[1] - b[1]).abs().max()), a[2]["kl"], b[2]["kl"], a[2]["lr"], b[2]["lr"])


if __name__ == "__main__":
    main()
