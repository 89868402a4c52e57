#!/usr/bin/env python
"""Train a task with the CUDA PPO engine for a number of iterations and print the learning curve (mean reward per rollout,
KL, learning rate) -- evidence that the env step, the observations and the PPO update form a loop that learns.
Usage: tools/train_curve.py [--task grasp_sim|orient|search] [--num-envs 4096] [--iters 300] [--minibatch 8192]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="grasp_sim")
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--minibatch", type=int, default=8192)
    ap.add_argument("--every", type=int, default=10)
    args = ap.parse_args()
    import torch
    from seqdex_b200.ppo import A2CAgent, PPOConfig
    from seqdex_b200.tasks import BlockAssemblyGraspSim, BlockAssemblyOrient, BlockAssemblySearch
    from seqdex_b200.vec_task import RLgamesVecTaskPython
    cls, ep, ema = {"grasp_sim": (BlockAssemblyGraspSim, 150, 1.0), "orient": (BlockAssemblyOrient, 75, 0.2),
                    "search": (BlockAssemblySearch, 75, 0.6)}[args.task]
    cfg = {"env": {"numEnvs": args.num_envs, "episodeLength": ep, "actionsMovingAverage": ema}, "sim": {"substeps": 2, "physx": {}},
           "task": {"randomize": False}}
    kw = {} if args.task == "search" else {"bank_per_type": 16}
    task = cls(cfg, **kw)
    agent = A2CAgent(RLgamesVecTaskPython(task, "cuda:0"), PPOConfig(minibatch_size=args.minibatch))
    t0 = time.time()
    acc, curve = [], []
    for it in range(args.iters):
        info = agent.train_epoch()
        acc.append(info["mean_reward"])
        if (it + 1) % args.every == 0:
            row = {"iter": it + 1, "env_steps": (it + 1) * agent.B, "mean_reward": sum(acc) / len(acc), "kl": info["kl"], "lr": info["lr"],
                   "wall_s": round(time.time() - t0, 1)}
            curve.append(row)
            print(json.dumps(row), flush=True)
            acc = []
    first, last = curve[0]["mean_reward"], curve[-1]["mean_reward"]
    print(json.dumps({"task": args.task, "num_envs": args.num_envs, "iters": args.iters, "first": first, "last": last,
                      "best": max(r["mean_reward"] for r in curve)}))


if __name__ == "__main__":
    main()
