#!/usr/bin/env python
"""Per-stage and per-warp cycle accounting of k_simulate (SIM_PROFILE build: clock64 of thread 0 after every stage barrier of a
sub-step, and per warp at both barriers of each of the 17 solver passes).
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared -DSIM_PROFILE \
       -o seqdex_b200/libseqdex_b200_prof.so seqdex_b200/csrc/sdx_env.cu seqdex_b200/csrc/sdx_ppo.cu
  SEQDEX_B200_LIB=$PWD/seqdex_b200/libseqdex_b200_prof.so python tools/sim_phase_cycles.py      # on a B200
Output of the round's runs: profiles/r01_sim_stage_cycles.txt, r01_ab_*.txt."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seqdex_b200.env import SdxEnv, make_heap_bank      # noqa: E402
from seqdex_b200.scene import Scene                      # noqa: E402
from seqdex_b200.tasks.block_assembly_grasp_sim import default_tvalue_weights   # noqa: E402

n = int(os.environ.get('SIM_PROF_ENVS', '2048'))
TASK = os.environ.get('SIM_PROF_TASK', 'BlockAssemblyGraspSim')        # or ToolPositioningGrasp / ToolPositioningOrient (one free body)
if TASK.startswith('ToolPositioning'):
    from seqdex_b200.tasks.cfg import scene_from_cfg
    from seqdex_b200.tasks.tool_positioning import synthetic_tool_grasp_bank
    scene = scene_from_cfg(TASK)
    env = SdxEnv(scene, n)
    if TASK.endswith('Orient'):
        env.set_grasp_bank(*synthetic_tool_grasp_bank(scene, 8, 0))
else:
    scene = Scene(edge_contacts=os.environ.get('SIM_PROF_EDGE', '1') != '0')
    bank = make_heap_bank(scene, 8)
    env = SdxEnv(scene, n)
    env.set_heap_bank(bank)
    env.set_tvalue_weights(default_tvalue_weights(22))
gen = torch.Generator(device="cuda").manual_seed(1)
env.step(torch.rand(n, 23, device="cuda", generator=gen) * 2 - 1)
env.tensor("PROGRESS").copy_(torch.randint(0, 75, (n,), device="cuda", generator=gen))
for _ in range(80):
    env.step(torch.rand(n, 23, device="cuda", generator=gen) * 2 - 1)
con = env.tensor("CONTACTS")                              # switches the dump (and the profile sink) on
env.step(torch.rand(n, 23, device="cuda", generator=gen) * 2 - 1)
torch.cuda.synchronize()
REC = 18 * 2 * 8 + 16                                     # SIM_PROF_REC in csrc/sdx_sim.cuh
ptr = con.data_ptr() + n * 1024 * 8 * 4
raw = torch.empty(n * 2 * REC, dtype=torch.int64, device="cuda")
ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(raw.data_ptr()), ctypes.c_void_p(ptr), n * 2 * REC * 8, 3)
p_all = raw.cpu().numpy().reshape(n * 2, REC).astype(np.float64)
p = p_all[p_all[:, 289] > 0]                                  # solver statistics: sub-steps that have contacts (the others skip it)
t = p[:, :18 * 2 * 8].reshape(-1, 18, 2, 8)[:, :17]           # [env x sub-step, pass (it = -1..15), {arrive after A, arrive after B}, warp]
t0, ncon, nact, nrob = p[:, 288], p[:, 289], p[:, 290], p[:, 291]
relA = t[:, :, 0, :].max(axis=2)                              # barrier release times = slowest arrival
relB = t[:, :, 1, :].max(axis=2)
start = np.concatenate([t0[:, None], relB[:, :-1]], axis=1)   # a pass starts when the previous pass's second barrier releases
durA, durB = relA - start, relB - relA
workA = t[:, :, 0, :] - start[:, :, None]                     # per warp: its own time in phase A / phase B
workB = t[:, :, 1, :] - relA[:, :, None]
marks = p_all[:, 292:304]
names = ["kinematics + free velocities + twists", "world AABBs", "broad phase", "pair offsets scan", "narrow pass 1 (pair masks)",
         "contact offsets", "narrow pass 2 (contacts + warm start)", "incidence lists + work items", "effective masses",
         "solver: 17 passes", "integrate + impulse cache"]
stage = np.diff(marks, axis=1)
tot = marks[:, 11] - marks[:, 0]
print(f"sub-step total {tot.mean():.0f} cycles (thread 0, barrier to barrier)")
for k, nm in enumerate(names):
    print(f"  {nm:40s} {stage[:, k].mean():8.0f} cycles  {100 * stage[:, k].mean() / tot.mean():5.1f} %")
print(f"sub-steps without a single contact (solver and incidence stages skipped): {100 * (p_all[:, 289] == 0).mean():.1f} %")
print(f"{len(p)} env x sub-step records; contacts {ncon.mean():.0f}, awake touched bricks {nact.mean():.1f}, robot links in contact {nrob.mean():.2f}")
print(f"solver loop {(relB[:, -1] - t0).mean():.0f} cycles per sub-step = 17 passes")
print(f"phase A per pass: {durA[:, 1:].mean():.0f} cycles (p90 {np.percentile(durA[:, 1:], 90):.0f});  phase B per pass: {durB.mean():.0f} (p90 {np.percentile(durB, 90):.0f})")
print("mean own work per pass by warp (cycles):")
print("   A:", " ".join(f"{workA[:, 1:, w].mean():6.0f}" for w in range(8)))
print("   B:", " ".join(f"{workB[:, :, w].mean():6.0f}" for w in range(8)))
slow = workB.argmax(axis=2).reshape(-1)
print("slowest warp of phase B:", np.round(np.bincount(slow, minlength=8) / slow.size, 3))
r = nrob > 0
print(f"robot in contact in {100 * r.mean():.0f} % of sub-steps: phase B {durB[r].mean():.0f} cycles/pass there (robot warp {workB[r][:, :, 7].mean():.0f}), {durB[~r].mean():.0f} elsewhere")
for lo, hi in ((0, 20), (20, 40), (40, 60), (60, 73)):
    m = (nact >= lo) & (nact < hi) & ~r
    if m.any():
        print(f"  awake touched bricks in [{lo},{hi}): {100 * m.mean():4.0f} % of records, phase B {durB[m].mean():5.0f} cycles/pass, phase A {durA[m][:, 1:].mean():5.0f}, contacts {ncon[m].mean():.0f}")
