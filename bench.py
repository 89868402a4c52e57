#!/usr/bin/env python
"""bench.py -- env-steps/sec of the BlockAssemblyGraspSim hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU baseline arm (oracle port, host cores)

A "step" is one VecTask.step() of every env on the rank: reset_idx + pre_physics (IK) + contact step +
observations/reward/t-value.  Weak scaling: NUM_ENVS envs per GPU.  Timing: CUDA events on the launching
stream, barrier + synchronize on both sides, max over ranks.  Inputs (260 MB of env state at 16384 envs)
exceed the 126 MB L2, so no explicit flush is needed between iterations (stated in config).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_ENV_STEP = 15956   # algorithmic HBM bytes / env-step (fp32 rollout)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_simulate launch at 16 384 envs, from the committed ncu --set full capture
# (profiles/r02_ncu_k_simulate_final2.txt: 180.2 MB read + 363.7 MB written; the writes include the contact records and the
# edge-contact normals that live in global memory behind L1 and the warm-start impulse cache)
NCU_TRAFFIC_BYTES_PER_ENV = (180.210e6 + 363.700e6) / 16384
NCU_ISSUE_ACTIVE = 0.4974            # smsp__issue_active.avg.pct_of_peak_sustained_active of the same capture
ALGO_FLOP_PER_ENV_STEP = 1.48e6      # counted fp32 work of the contact step in this episode mix (oracle counters)
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # B200: 148 SMs x 128 fp32 lanes x 2 (FMA) x 1.965 GHz


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def measured_tensor_peak():
    """dense bf16 TFLOP/s of this pool's B200s: the SUSTAINED figure (the GEMMs run inside a long step), else the recipe's fallback"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md)"


# algorithmic tensor work of PPO per env-step (with the a2c net's dead critic trunk left out):
#   inference : actor 396-1024-512-256-23 + central value 564-1024-512-256-1, one forward each per env-step
#   update    : 5 mini-epochs x (forward + dX + dW) of both nets over every sample
ACTOR_MAC = 396 * 1024 + 1024 * 512 + 512 * 256 + 256 * 23
CV_MAC = 564 * 1024 + 1024 * 512 + 512 * 256 + 256 * 1
PPO_FLOP_PER_ENV_STEP = 2.0 * (ACTOR_MAC + CV_MAC) * (1 + 5 * 3)


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


PRE_STEPS = 80          # untimed steps after staggering: every env has been through a reset, the episode-phase mix is stationary
STAGGER = 75            # an untrained policy's episodes end at progress 75 (GS:1751: far from the target after 75 steps)


def precondition_cpu(env, rng, stagger=True):
    """same episode-phase mix as the GPU arm: first step resets every env, then progress ~ U[0, 75), then PRE_STEPS steps"""
    n = env.n
    env.step(rng.uniform(-1, 1, size=(n, 23)).astype(np.float32))
    if stagger:
        env.progress[:] = rng.integers(0, STAGGER, size=n)
    for _ in range(PRE_STEPS):
        env.step(rng.uniform(-1, 1, size=(n, 23)).astype(np.float32))


def cpu_baseline(scene, seconds=12.0, n_envs=None, bank=None, stagger=True):
    """the oracle (CPU port of the same hot path) on the host cores, bounded sample of the same workload"""
    from oracle import oracle
    oracle.build()
    cores = oracle.lib().sdxo_get_threads()
    n = n_envs or 32 * cores
    env = oracle.OracleEnv(scene, n)
    if bank is not None and int(scene.c.task) in (3, 5):
        env.set_grasp_bank(*bank)
    elif bank is not None:
        env.set_heap_bank(bank)
    if int(scene.c.task) == 2:
        from seqdex_b200.camera import SEARCH_CAMERA, look_at
        env.set_camera(look_at(**SEARCH_CAMERA))
    rng = np.random.default_rng(0)
    precondition_cpu(env, rng, stagger)
    t0, steps = time.time(), 0
    while time.time() - t0 < seconds or steps < 2:
        env.step(rng.uniform(-1, 1, size=(n, 23)).astype(np.float32))
        steps += 1
    dt = time.time() - t0
    return {"value": n * steps / dt, "unit": "env-steps/s", "cores": int(cores), "kind": "port",
            "sample": f"{n} envs x {steps} VecTask.step() calls of the C oracle (oracle/sdx_oracle.c), {dt:.1f} s, after the same "
                      f"episode staggering + {PRE_STEPS} untimed steps as the GPU arm"}


def host_bank(scene, per_type=2, settle=150):
    """small settled heap bank produced by the ORACLE (used only by the CPU arms)"""
    from oracle import oracle
    oracle.build()
    n = 8 * per_type
    old = scene.c.brick_lin_damp
    scene.c.brick_lin_damp = 10.0
    env = oracle.OracleEnv(scene, n)
    rng = np.random.default_rng(1)
    env.brick[:, 0:2, :] += rng.uniform(-0.01, 0.01, size=(n, 2, 72)).astype(np.float32)
    for _ in range(settle):
        env.simulate()
    scene.c.brick_lin_damp = old
    for _ in range(settle // 3):
        env.simulate()
    rows = env.brick_roots()
    rows[..., 7:13] = 0
    return rows.reshape(8, per_type, 72, 13)


TASKS = {"grasp_sim": ("BlockAssemblyGraspSim", 396), "orient": ("BlockAssemblyOrient", 186), "search": ("BlockAssemblySearch", 186),
         "insert": ("BlockAssemblyInsertSim", 75), "tool_grasp": ("ToolPositioningGrasp", 468), "tool_orient": ("ToolPositioningOrient", 468)}


def make_config(args, world):
    """the `config` object of the JSON line -- identical for this arm and for --impl reference (same workload, two implementations)"""
    task_name = TASKS[args.task][0]
    n = args.num_envs
    lockstep = args.task in ("orient", "search")
    mbs = min(args.minibatch, 8 * n)
    return {"workload": (f"{task_name} num_envs={n} per GPU, PPO bf16: every 8 env steps (policy + central-value forward, "
                         "VecTask.step = reset_idx + IK + contact step 2x16 + obs/reward/t-value) then GAE and "
                         f"5 mini-epochs x {8 * n // mbs} minibatches for actor and central value"
                         if args.mode == "ppo" else
                         f"{task_name} num_envs={n} per GPU, rollout only: VecTask.step, U(-1,1) actions"),
            "mode": args.mode, "minibatch": mbs,
            "num_envs_per_gpu": n, "global_envs": n * world, "parallelism": f"env-sharded x{world}, no data-path collective",
            "l2": "env state (260 MB at 16384 envs) exceeds the 126 MB L2; no explicit flush",
            "heap_bank_per_type": args.bank_per_type,
            "episodes": (f"lockstep as in the reference (time-outs every 75 steps; each reset runs the scripted reset_idx inside the "
                         f"timed region: 103 contact steps for Orient OR:1390-1695, 60 + render + 1 + render for Search "
                         f"SE:1274-1537, 989-1019); {PRE_STEPS} untimed steps before warm-up" if lockstep else
                         f"staggered: progress ~ U[0,{STAGGER}) then {PRE_STEPS} untimed steps before warm-up (stationary mix of "
                         "fresh and settled heaps; resting bricks sleep with PhysX's default threshold and 0.4 s timer)")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from seqdex_b200.scene import Scene
    from oracle import oracle
    scene = Scene()
    oracle.build()
    cores = oracle.lib().sdxo_get_threads()
    n = 32 * cores
    bank = host_bank(scene)
    env = oracle.OracleEnv(scene, n)
    env.set_heap_bank(bank)
    rng = np.random.default_rng(0)
    precondition_cpu(env, rng)
    for _ in range(max(args.warmup, 1)):
        env.step(rng.uniform(-1, 1, size=(n, 23)).astype(np.float32))
    t0 = time.time()
    for _ in range(args.steps):
        env.step(rng.uniform(-1, 1, size=(n, 23)).astype(np.float32))
    dt = time.time() - t0
    val = n * args.steps / dt
    sample = (f"{n} envs per step (bounded sample of the {args.num_envs}-env workload), C oracle on {cores} host threads, "
              f"episodes staggered + {PRE_STEPS} untimed steps first")
    emit({
        "impl": "reference", "metric": "env-steps/sec at num_envs=16384 (BlockAssemblyGraspSim)", "value": val, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args, max(args.gpus, 1)),
        "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": int(cores), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Isaac Gym (closed binary) is not installable here; this arm times the repo's CPU oracle of the same path: VecTask.step of "
                f"a {n}-env sample (the rollout; PPO's tensor work is not part of the CPU arm)",
    })


def run_chain_bench(args):
    """BASELINE configs[3]: the task chain with transition-feasibility (t-value) gates, every rank on its own shard of envs with the
    hand-offs resident on the device (seqdex_b200/chain.py).  `--num-envs` is PER GPU: configs[3] = --gpus 8 --num-envs 1024."""
    import torch
    import torch.distributed as dist
    from seqdex_b200.chain import run_chain
    from seqdex_b200.tasks.block_assembly_grasp_sim import default_tvalue_weights
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.num_envs
    w = default_tvalue_weights(1)
    w[-1] += 50.0            # no trained gate checkpoints ship with the reference: the success logit is biased open so every stage hands on
    eps = max(1, -(-args.steps // 75))
    sampler = ClockSampler(local)
    sampler.start()
    timing = {}
    out = run_chain(num_envs=n, device_id=local, episodes=(eps, eps, eps, eps), tvalue_weights=w, bank_capacity=max(64, n // 8), seed=22 + rank,
                    timing=timing)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    stages = list(timing)
    t = torch.tensor([timing[k][0] for k in stages], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs = dict(zip(stages, t.tolist()))
    steps = {k: timing[k][1] for k in stages}
    tot_s, tot_steps = sum(secs.values()), sum(steps.values())
    if rank == 0:
        emit({
            "metric": "env-steps/sec, full chain with t-value switching (BASELINE configs[3])", "value": world * n * tot_steps / tot_s,
            "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * tot_s / tot_steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"chain {' -> '.join(stages)} at num_envs={n} per GPU ({n * world} global), {eps} episode(s) per stage, uniform-random "
                                   "policies, device-resident hand-offs (heap rings -> next stage's bank), gates = the stages' t-value / pixel tests",
                       "num_envs_per_gpu": n, "global_envs": n * world, "parallelism": f"env-sharded x{world}, no data-path collective",
                       "timed": "VecTask.step loops of every stage incl. the scripted resets (CUDA events, max over ranks); env construction and bank hand-over outside"},
            "per_stage": {k: {"env_steps_per_s": world * n * steps[k] / secs[k], "steps": steps[k], "seconds": secs[k]} for k in stages},
            "gates": {"search_heaps_per_type": out.get("search_heaps_per_type"), "orient_heaps_per_type": out.get("orient_heaps_per_type"),
                      "grasp_terminal_states": out.get("grasp_terminal_states"), "insert_success_rate": out.get("insert_success_rate")},
            "mean_rewards": {k: out.get(f"{k}_mean_reward") for k in stages},
            "e2e": {"value": world * n * tot_steps / tot_s, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "scope": "device-resident chain (policies on the device); the host-buffer path is measured by --task grasp_sim"},
            "clocks": sampler.summary(),
        })
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """the contract is ONE JSON line on stdout: anything libraries print there (NCCL's version banner, torchrun notices) is
    routed to stderr by pointing fd 1 at fd 2; emit() writes the line to the saved descriptor"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    line = json.dumps(obj)
    if _REAL_STDOUT is not None:
        _REAL_STDOUT.write(line + "\n")
        _REAL_STDOUT.flush()
    else:
        print(line)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=160)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--num-envs", type=int, default=16384, help="envs PER GPU (weak scaling)")
    ap.add_argument("--bank-per-type", type=int, default=5000,
                    help="settled heaps per brick type in the bank reset_idx samples: the reference samples rows [0, 5000) of its pickle (GS:1508)")
    ap.add_argument("--e2e-steps", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sleep-off", action="store_true", help="skip the extra rollout with sleeping switched off (reported beside the default)")
    ap.add_argument("--edge-contacts", type=int, default=1, choices=[0, 1],
                    help="0: the contact step without edge-edge contacts (the round-1 / early round-2 geometry; DESIGN.md section 3c), for A/B runs")
    ap.add_argument("--mode", default="ppo", choices=["ppo", "rollout"],
                    help="ppo: PPO training in the loop (policy forward, env step, GAE, 5 mini-epochs of updates every 8 steps); "
                         "rollout: VecTask.step only with U(-1,1) actions")
    ap.add_argument("--task", default="grasp_sim", choices=["grasp_sim", "orient", "search", "insert", "chain", "tool_grasp", "tool_orient"],
                    help="grasp_sim: BlockAssemblyGraspSim, the BASELINE.json metric (configs[1]); orient: BlockAssemblyOrient (configs[2]: "
                         "32768 envs over 2 GPUs = --gpus 2 with the default 16384 envs per GPU); tool_grasp / tool_orient: the two "
                         "ToolPositioning tasks of configs[4] (65536 envs over 8 GPUs = --gpus 8 --num-envs 8192)")
    ap.add_argument("--minibatch", type=int, default=32768,
                    help="PPO minibatch; the yaml's 4 is a 4-env smoke value (SURVEY.md section 7): default = batch/4 at 16384 envs x horizon 8")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.task == "chain":
        return run_chain_bench(args)

    import torch
    import torch.distributed as dist
    from seqdex_b200.env import SdxEnv, make_heap_bank
    from seqdex_b200.scene import Scene
    from seqdex_b200.tasks.block_assembly_grasp_sim import default_tvalue_weights
    from seqdex_b200.tasks.cfg import scene_from_cfg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = args.num_envs
    orient = args.task in ("orient", "search")      # lockstep episodes (both tasks' resets are scripts over the whole sim)
    search = args.task == "search"
    # Orient's reset_idx is a script over the WHOLE sim (103 extra contact steps whenever any env resets, OR:1390-1695), so its
    # episodes run in lockstep as in the reference (time-outs only); GraspSim's per-env resets are staggered (see below)
    insert = args.task == "insert"
    task_name = TASKS[args.task][0]
    scene = scene_from_cfg(task_name, edge_contacts=bool(args.edge_contacts))   # the task's yaml-stated sim / env parameters (contact_offset 0.02 for Orient / Search)
    obs_dim, state_dim = TASKS[args.task][1], (188 if insert else 564)
    env = SdxEnv(scene, n, local, seed=22 + rank)
    if args.task == "tool_grasp":                   # resets to a fixed start pose (TG:1459-1578): no bank
        bank = None
    elif args.task == "tool_orient":                # restores banked grasps (TO:365-368): synthetic stand-ins here
        from seqdex_b200.tasks.tool_positioning import synthetic_tool_grasp_bank
        bank = synthetic_tool_grasp_bank(scene, min(args.bank_per_type, 64), seed=22 + rank)
        env.set_grasp_bank(*bank)
    elif insert:                                    # InsertSim restores banked grasps (IS:372-375): synthetic stand-ins here, GraspSim's rings in --task chain
        from seqdex_b200.tasks.block_assembly_insert_sim import synthetic_grasp_bank
        bank = synthetic_grasp_bank(scene, min(args.bank_per_type, 64), seed=22 + rank)
        env.set_grasp_bank(*bank)
    elif search:                                      # Search resets from the drop lattice and renders its overview camera
        from seqdex_b200.camera import SEARCH_CAMERA, look_at
        env.set_camera(look_at(**SEARCH_CAMERA))
        bank = None
    else:
        bank = make_heap_bank(scene, args.bank_per_type, local, seed=22 + rank)
        env.set_heap_bank(bank)
    env.set_tvalue_weights(default_tvalue_weights(22))
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    K, W = args.steps, max(args.warmup, 3)
    # pre-generated U(-1,1) actions (SURVEY 8d input B), resident in HBM before the timed region
    acts = torch.rand(K + W, n, 23, device=dev, generator=gen) * 2 - 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import ctypes
    from seqdex_b200 import _lib
    ppo_launch = lambda: int(_lib.load().sdx_ppo_launch_count())
    # stationary episode-phase mix (untimed set-up): the first step resets every env, then progress ~ U[0, 75) and
    # PRE_STEPS steps, so that resets -- and with them the waking / falling asleep of the heaps -- are spread over time
    # as they are in a long training run instead of all envs marching through one episode in lockstep
    env.step(acts[0])
    if not orient:
        env.tensor("PROGRESS").copy_(torch.randint(0, STAGGER, (n,), device=dev, generator=gen))
    for i in range(PRE_STEPS):
        env.step(torch.rand(n, 23, device=dev, generator=gen) * 2 - 1)
    for i in range(W):
        env.step(acts[i])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    # ---- (1) rollout only: VecTask.step with actions resident in HBM; per-launch events around the contact step
    KR = min(K, 64)
    l0 = env.launch_count()
    sim_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KR)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(KR):
        env.pre_physics(acts[W + i])
        sim_ev[i][0].record()
        env.simulate()
        sim_ev[i][1].record()
        env.post_physics()
    e1.record()
    barrier()
    ro_ms = e0.elapsed_time(e1)
    ro_launches = env.launch_count() - l0
    sim_ms = float(np.mean([a.elapsed_time(b) for a, b in sim_ev]))
    ms, launches, ppo_info = ro_ms * K / KR, ro_launches * K // KR, None
    if args.mode == "ppo":
        # ---- (2) PPO in the loop: the headline.  K env steps = K/8 iterations of play_steps + train_epoch
        from seqdex_b200.ppo import A2CAgent, PPOConfig
        from seqdex_b200.vec_task import RLgamesVecTaskPython

        class _Task:      # the task surface VecTask needs, over the SAME env (no second simulation state)
            pass
        task = _Task()
        task.env, task.num_envs, task.num_obs, task.num_states, task.num_actions, task.device = env, n, obs_dim, state_dim, 23, f"cuda:{local}"
        task.obs_buf, task.states_buf, task.rew_buf, task.reset_buf = (env.tensor(k) for k in ("OBS", "STATES", "REW", "RESET"))
        task.extras = {}
        step_ev = []                    # CUDA events around every VecTask.step of the timed PPO iterations
        def _timed_step(a):
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(); r = env.step(a); ev[1].record()
            step_ev.append(ev)
            return r
        task.step = _timed_step
        venv = RLgamesVecTaskPython(task, task.device)
        agent = A2CAgent(venv, PPOConfig(minibatch_size=min(args.minibatch, 8 * n)), device=local,
                         dist_group=dist.group.WORLD if world > 1 else None)
        H = agent.H
        # whole PPO iterations covering at least K env steps; at least TWO untimed iterations first: the agent's second update() call
        # captures the update into a CUDA graph (ppo.A2CAgent.update), and one-off set-up does not belong in the timed region
        iters, witers = max(-(-K // H), 1), max(-(-W // H), 2)
        for _ in range(witers):
            ppo_info = agent.train_epoch()
        barrier()
        l0, p0 = env.launch_count(), ppo_launch()
        step_ev.clear()
        asleep0 = float((env.tensor("SLEEP") >= max(scene.c.sleep_substeps, 1)).float().mean())
        contacts0 = float(env.tensor("NCONTACT")[:, 0].float().mean())
        it_ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
        e0.record()
        it_ev[0].record()
        for it in range(iters):
            ppo_info = agent.train_epoch()
            it_ev[it + 1].record()
        e1.record()
        barrier()
        timed_s = e0.elapsed_time(e1) * 1e-3
        ms = e0.elapsed_time(e1) * K / (iters * H)
        it_ms = [it_ev[i].elapsed_time(it_ev[i + 1]) for i in range(iters)]
        ppo_info = dict(ppo_info or {})
        ppo_info["timed_region"] = {"seconds": timed_s, "env_steps_timed": iters * H, "ppo_iterations": iters,
                                    "ms_per_iteration_min_median_max": [float(np.min(it_ms)), float(np.median(it_ms)), float(np.max(it_ms))],
                                    "ms_per_iteration_first_last": [float(it_ms[0]), float(it_ms[-1])],
                                    "contacts_per_env_mean_start_end": [contacts0, float(env.tensor("NCONTACT")[:, 0].float().mean())],
                                    "drift": "the policy is being TRAINED inside the timed region: as it learns to reach for the target brick the hand "
                                             "spends more time in the heap, contacts per env and with them the contact step's time rise over the "
                                             "iterations (a longer --steps therefore reports a lower value than a short one)",
                                    "note": f"`value` = envs x {iters * H} steps / that time (whole iterations covering the {K} steps asked for)"}
        from seqdex_b200.dist_utils import params_digest
        dg = params_digest(agent.actor.params, agent.cv.params)
        if world > 1:
            dgs = [torch.zeros_like(dg) for _ in range(world)]
            dist.all_gather(dgs, dg)
        else:
            dgs = [dg]
        ppo_info["lockstep"] = {"params_digest": [int(x) for x in dg.tolist()], "ranks": world,
                                "all_ranks_equal": bool(all(torch.equal(dgs[0], x) for x in dgs)),
                                "note": "digest of actor + central-value parameters all-gathered after the timed region: data-parallel replicas must be bit-identical"}
        ppo_info["env_step_ms_in_loop"] = float(np.mean([a.elapsed_time(b) for a, b in step_ev]))
        ppo_info["ppo_ms_per_iteration"] = float(np.median(it_ms)) - H * ppo_info["env_step_ms_in_loop"]
        ppo_info["bricks_asleep_frac_start_end"] = [asleep0, float((env.tensor("SLEEP") >= max(scene.c.sleep_substeps, 1)).float().mean())]
        launches = (env.launch_count() - l0 + ppo_launch() - p0) * K // (iters * H)
    # ---- end to end through the C-ABI with HOST buffers (sdx_step_host): H2D actions, D2H obs/states/rew/reset
    E = 75 if orient and args.e2e_steps >= 32 else args.e2e_steps   # Orient: one whole episode, so one scripted reset is inside
    h_act = torch.empty(n, 23, dtype=torch.float32).pin_memory()
    h_act.copy_(acts[0].cpu())
    h_obs = torch.empty(n, obs_dim, dtype=torch.float32).pin_memory()
    h_st = torch.empty(n, state_dim, dtype=torch.float32).pin_memory()
    h_rew = torch.empty(n, dtype=torch.float32).pin_memory()
    h_rs = torch.empty(n, dtype=torch.int64).pin_memory()
    for _ in range(3):
        env.step_host(h_act, h_obs, h_st, h_rew, h_rs)
    barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(E):
        env.step_host(h_act, h_obs, h_st, h_rew, h_rs)
    g1.record()
    barrier()
    e2e_wall_ms = 1000 * (time.perf_counter() - t0)      # host clock around the same region (sdx_step_host ends in a stream synchronize)
    e2e_ms = g0.elapsed_time(g1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # ---- what the solver costs on a LIVE heap: the same rollout with sleeping switched off (PhysX's sleeping is on by default and so is
    #      ours; this number is reported beside the default so the reader sees what the mechanism saves) -- untimed set-up, rank 0 only
    sleep_off = edge_off = None

    def companion_rollout(note, own_bank=False, **scene_kw):
        scene2 = scene_from_cfg(task_name, **scene_kw)
        # a heap banked under one contact model is not at rest under another (a brick that rested on an edge-edge contact starts to sink and
        # wakes its neighbours): a companion that changes the contact model settles its own bank, exactly as the main arm did
        bank2 = make_heap_bank(scene2, args.bank_per_type, local, seed=22 + rank) if own_bank else bank
        env2 = SdxEnv(scene2, n, local, seed=22 + rank)
        env2.set_heap_bank(bank2)
        env2.set_tvalue_weights(default_tvalue_weights(22))
        env2.step(acts[0])
        env2.tensor("PROGRESS").copy_(torch.randint(0, STAGGER, (n,), device=dev, generator=gen))
        for i in range(PRE_STEPS):
            env2.step(torch.rand(n, 23, device=dev, generator=gen) * 2 - 1)
        for i in range(W):                                                  # the same timeline as the main arm's rollout-only phase
            env2.step(acts[i])
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KR)]
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(KR):
            env2.pre_physics(acts[W + i])
            ev2[i][0].record(); env2.simulate(); ev2[i][1].record()
            env2.post_physics()
        s1.record()
        torch.cuda.synchronize()
        nc2 = env2.tensor("NCONTACT").cpu().numpy()
        res = {"k_simulate_ms_per_launch": float(np.mean([a.elapsed_time(b) for a, b in ev2])),
               "rollout_env_steps_per_s": n * KR / (s0.elapsed_time(s1) * 1e-3), "steps": KR, "contacts_per_env_mean": float(nc2[:, 0].mean()),
               "envs_shedding_frac": float((nc2[:, 2] > 0).mean()), "dropped_max": int(nc2[:, 1].max()),
               "bricks_asleep_frac": float((env2.tensor("SLEEP") >= scene2.c.sleep_substeps).float().mean()) if scene2.c.sleep_substeps else 0.0,
               "note": note}
        env2.close()
        return res

    if rank == 0 and args.task == "grasp_sim" and not args.no_sleep_off:
        sleep_off = companion_rollout("rank 0, same episode mix and bank, Scene(sleep_time=0): no brick ever sleeps",
                                      sleep_time=0.0, edge_contacts=bool(args.edge_contacts))
        # what the edge-edge contacts cost (round 1 / early round 2 ran without them): the same rollout with the
        # instantiation of k_simulate that carries none of their code
        if args.edge_contacts:
            edge_off = companion_rollout("rank 0, same episode mix and timeline, Scene(edge_contacts=False): corner-vs-face contacts only "
                                         "(k_simulate<.., EDGE = false>) on a bank settled under that model, as every measurement before the last week "
                                         "of round 2; compare with rollout_only / roofline.ms_per_launch", own_bank=True, edge_contacts=False)
    tms = torch.tensor([ms, e2e_ms, sim_ms, ro_ms, e2e_wall_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms, sim_ms, ro_ms, e2e_wall_ms = (float(x) for x in tms.tolist())
    nc = env.tensor("NCONTACT").cpu().numpy()
    if rank == 0:
        peak, which = measured_peaks()
        achieved = ALGO_BYTES_PER_ENV_STEP * n / (sim_ms * 1e-3) / 1e9
        out = {
            "metric": f"env-steps/sec at num_envs={n} ({task_name})" if n != 16384 or args.task != "grasp_sim" else "env-steps/sec at num_envs=16384 (BlockAssemblyGraspSim)", "value": world * n * K / (ms * 1e-3),
            "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args, world),
            "e2e": {"value": world * n * E / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": n * 23 * 4,
                    "d2h_bytes_per_step": n * (obs_dim + state_dim + 1) * 4 + n * 8, "steps": E, "timer": "cuda events (max over ranks)",
                    "value_wall_clock": world * n * E / (e2e_wall_ms * 1e-3),
                    "scope": "VecTask.step through the C-ABI with pinned HOST buffers (sdx_step_host): the rollout only -- PPO is NOT in this "
                             "loop, which is why it can exceed `value` (PPO in the loop, device-resident)"},
            "gpu_launches": int(launches),
            "rollout_only": {"value": world * n * KR / (ro_ms * 1e-3), "unit": "env-steps/s", "steps": KR, "ms_per_step": ro_ms / KR},
            "ppo": ppo_info,
            "roofline": {"bound": "hbm", "kernel": "k_simulate", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)", "traffic": NCU_TRAFFIC_BYTES_PER_ENV * n,
                         "traffic_source": "ncu --set full at 16384 envs, profiles/r02_ncu_k_simulate_final2.txt (scaled by envs per launch)",
                         "issue_slots_busy_ncu": NCU_ISSUE_ACTIVE,
                         "fp32_alu": {"algorithmic_flop_per_env_step": ALGO_FLOP_PER_ENV_STEP, "peak_tflops": FP32_PEAK_TFLOPS,
                                      "achieved_tflops": ALGO_FLOP_PER_ENV_STEP * n / (sim_ms * 1e-3) / 1e12,
                                      "frac": ALGO_FLOP_PER_ENV_STEP * n / (sim_ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS,
                                      "note": "GraspSim mix; neither roofline bounds the kernel (DESIGN.md sections 6, 11)"},
                         "ms_per_launch": sim_ms, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * n,
                         "share_of_step": sim_ms * K / ms, "share_of_rollout_step": sim_ms * KR / ro_ms,
                         "note": "state-streaming bound is loose: the kernel is instruction- and latency-bound (2.7 G warp-instructions per launch = 2.3 ms "
                                 "at full issue rate; issue slots 50 % busy, 41 % of the stall samples are block-barrier waits behind the slowest warp "
                                 "of a stage; DESIGN.md sections 11, 11.4), not HBM-bound"},
            "roofline_tensor": None,
            "clocks": sampler.summary(),
            "contacts_per_env": {"mean": float(nc[:, 0].mean()), "max": int(nc[:, 0].max()), "table": 1024,
                                 "dropped_max": int(nc[:, 1].max()),                       # contacts beyond the table AFTER shedding speculative ones
                                 "shed_level_max": int(nc[:, 2].max()), "envs_shedding_frac": float((nc[:, 2] > 0).mean()),
                                 "candidate_pairs_dropped_max": int((nc[:, 3] & 0xFFFF).max()),
                                 "envs_dropping_candidates_frac": float(((nc[:, 3] & 0xFFFF) > 0).mean()),
                                 "static_pairs_dropped_max": int((nc[:, 3] >> 16).max()),
                                 "contact_offset": float(scene.c.contact_offset),
                                 "note": "last step of the run, over all envs of rank 0; shed level 1/2/3 = speculative range halved / quartered / "
                                         "touching contacts only (csrc/sdx_sim.cuh); statics claim candidate slots first"},
            "bricks_asleep_frac": float((env.tensor("SLEEP") >= scene.c.sleep_substeps).float().mean()) if scene.c.sleep_substeps else 0.0,
            "sleep_off": sleep_off,
            "edge_off": edge_off,
        }
        if args.mode == "ppo" and args.task == "grasp_sim" and ppo_info.get("ppo_ms_per_iteration", 0) > 0:
            tpeak, tsrc = measured_tensor_peak()
            ach = PPO_FLOP_PER_ENV_STEP * n * 8 / (ppo_info["ppo_ms_per_iteration"] * 1e-3) / 1e12
            out["roofline_tensor"] = {
                "bound": "tensor", "kernels": "k_gemm_tn / k_gemm_tn2 (tcgen05, csrc/sdx_gemm.cuh) and everything else of the PPO part of an iteration",
                "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak, "peak_source": tsrc,
                "algorithmic_flop_per_env_step": PPO_FLOP_PER_ENV_STEP, "ppo_ms_per_iteration": ppo_info["ppo_ms_per_iteration"],
                "note": "algorithmic bf16 FLOPs of the policy / central-value forward passes and the 5 mini-epochs of updates over ALL PPO time of an "
                        "iteration (iteration time minus its 8 env steps): GEMMs, their conversions, loss / Adam kernels and launch gaps included; "
                        "per-kernel tensor-pipe counters: profiles/r02_ncu_gemm.txt"}
        if args.task.startswith("tool"):
            out["roofline"]["note"] = ("ToolPositioning scene (ONE free body of two boxes + the arm and hand): the per-env byte / flop / ncu constants above "
                                       "are the 72-brick GraspSim scene's and overstate this scene's traffic; ms_per_launch and the shares are measured")
        if not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(scene, bank=None if bank is None else bank if isinstance(bank, tuple) else bank.cpu().numpy(), stagger=not orient,
                                               n_envs=1024 if orient else None)
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
