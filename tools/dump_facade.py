#!/usr/bin/env python
"""The Isaac-Gym-shaped tensors of the facade (sdx_refresh: actor root states, rigid-body states, DoF states) after a few GraspSim steps on
the GPU, together with what the fused kernels computed from the SAME state (observations, privileged states, reward, reset flags, gate).
tests/test_facade_reference_cpu.py feeds the tensors to the REFERENCE's own compute_observations / compute_reward (in the build container,
where /root/reference exists) and expects the kernels' numbers: the claim of INTEGRATION.md section 2 -- the reference task's Python runs
unchanged on this facade -- as a test.  Run on a B200:  python tools/dump_facade.py   -> tests/golden/facade_dump.npz
(tests/test_facade_gpu.py re-creates the dump and compares it with the committed file bit for bit)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
N, STEPS = 8, 14


def make_dump():
    from seqdex_b200.env import SdxEnv
    from seqdex_b200.tasks.block_assembly_grasp_sim import default_tvalue_weights
    from seqdex_b200.tasks.cfg import scene_from_cfg
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import lattice_bank
    scene = scene_from_cfg("BlockAssemblyGraspSim")
    env = SdxEnv(scene, N, 0, seed=5)
    env.set_heap_bank(lattice_bank(scene, 2, seed=3))
    w = default_tvalue_weights(4)
    env.set_tvalue_weights(w)
    g = torch.Generator(device="cuda").manual_seed(9)
    prev_obs = prev_states = None
    for k in range(STEPS):
        prev_obs, prev_states = env.tensor("OBS").clone(), env.tensor("STATES").clone()
        a = torch.rand(N, 23, device="cuda", generator=g) * 2 - 1
        if k == STEPS - 1:
            env.tensor("PROGRESS")[0] = 148          # one env times out in the dumped step (GS:1745)
        env.step(a)
    for name in ("ROOT", "RB", "DOF_STATE"):
        env.refresh(name)
    torch.cuda.synchronize()
    t = lambda k: env.tensor(k).cpu().numpy().copy()
    d = dict(root=t("ROOT"), rb=t("RB").reshape(N, 165, 13), dof_state=t("DOF_STATE").reshape(N, 23, 2), netf=t("NETF"), actions=t("ACTIONS"),
             progress=t("PROGRESS"), target_init=t("TARGET_INIT"), obs=t("OBS"), states=t("STATES"), rew=t("REW"), reset=t("RESET"),
             tvalue=t("TVALUE"), consec=t("CONSEC"), prev_obs=prev_obs.cpu().numpy(), prev_states=prev_states.cpu().numpy(), tv_weights=w)
    env.close()
    return d


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden", "facade_dump.npz")
    if len(sys.argv) > 1:
        out = sys.argv[1]
    np.savez(out, **make_dump())
    print("wrote", out)
