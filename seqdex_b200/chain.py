"""The task chain Search -> Orient -> GraspSim (BASELINE ``configs[3]``, the links whose scenes are boxes; InsertSim's stud-level
insertion is outside this engine's contact model) with the reference's hand-offs kept on the device.

In the reference each stage is a separate ``train_rlgames.py`` run that ends by pickling what the next one loads:
    Search   writes  intermediate_state/saved_searching_ternimal_states_medium_mo_tvalue.pkl   (SE:1348-1352)
    Orient   reads it (OR:419-420), writes  saved_searching_ternimal_states_good_mo_tvalue.pkl (OR:1510-1512)
    GraspSim reads it (GS:412-413), writes  saved_grasping_{object,hand}_ternimal_states_*.pkl (GS:1448-1451)  -> InsertSim (IS:372-375)
Here a stage hands its device-resident rings straight to the next (``bank_io.*_bank_valid``); ``save_dir`` additionally writes the
same pickles, so a stage can also be resumed by -- or hand over to -- the reference's own scripts.

A stage's *transition-feasibility gate* decides what is worth handing on: enough pixels of the target brick visible (Search,
SE:1290, 1310), the brick face up as judged by the t-value network (Orient, OR:1203-1205, 1471), a lifted grasp the gate accepts
(GraspSim, GS:1402-1406).  ``policy(obs) -> actions`` drives a stage; the default is the uniform-random policy the benchmarks use.
"""
from __future__ import annotations

import os

import torch

from . import bank_io
from .tasks import BlockAssemblyGraspSim, BlockAssemblyOrient, BlockAssemblySearch
from .vec_task import RLgamesVecTaskPython


def _cfg(num_envs, episode, ema):
    return {"env": {"numEnvs": num_envs, "episodeLength": episode, "actionsMovingAverage": ema},
            "sim": {"substeps": 2, "physx": {}}, "task": {"randomize": False}}


def _random_policy(num_envs, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    return lambda obs: torch.rand(num_envs, 23, device=device, generator=g) * 2 - 1


def _run(env, policy, steps):
    obs = env.reset()
    rew_sum = torch.zeros((), device=obs["obs"].device)
    for _ in range(steps):
        obs, rew, _, _ = env.step(policy(obs["obs"]))
        rew_sum += rew.mean()
    return float(rew_sum) / max(steps, 1)


def run_chain(num_envs=256, device_id=0, episodes=(2, 2, 1), policies=None, tvalue_weights=None, bank_capacity=64, seed=22,
              save_dir=None, min_bank=1):
    """Run the three stages back to back on ``num_envs`` envs of GPU ``device_id``.

    episodes        -- episodes per stage (Search and Orient episodes are 75 steps, GraspSim's 150)
    policies        -- optional dict {'search' | 'orient' | 'grasp': callable(obs) -> actions}; default uniform random
    tvalue_weights  -- GraspInsertTValue parameters (flat, state_dict order) for the Orient / GraspSim gates
    Returns a dict of per-stage statistics and the three banks (device tensors)."""
    policies = policies or {}
    device = f"cuda:{device_id}"
    out = {}
    torch.manual_seed(seed)                         # train_rlgames.py:65 set_seed: VecTask.reset draws from the global generator
    # ---- stage 1: dig the target brick out of the heap
    search = BlockAssemblySearch(_cfg(num_envs, 75, 0.6), device_id=device_id, seed=seed, record_heaps=bank_capacity)
    env = RLgamesVecTaskPython(search, device)
    out["search_mean_reward"] = _run(env, policies.get("search") or _random_policy(num_envs, device, seed), 75 * episodes[0] + 1)
    heaps = bank_io.search_bank_valid(search.env)
    out["search_heaps_per_type"] = int(heaps.shape[1])
    if heaps.shape[1] < min_bank:
        raise RuntimeError("Search banked too few heaps to hand on")
    if save_dir:
        os.makedirs(save_dir, exist_ok=True)
        bank_io.save_search_bank(search.env, search.scene, os.path.join(save_dir, "saved_searching_ternimal_states_medium_mo_tvalue.pkl"),
                                 os.path.join(save_dir, "saved_searching_hand_ternimal_states_medium_mo_tvalue.pkl"))
    search.env.close()
    # ---- stage 2: turn it face up
    orient = BlockAssemblyOrient(_cfg(num_envs, 75, 0.2), device_id=device_id, seed=seed, heap_bank=heaps, tvalue_weights=tvalue_weights,
                                 record_heaps=bank_capacity)
    env = RLgamesVecTaskPython(orient, device)
    out["orient_mean_reward"] = _run(env, policies.get("orient") or _random_policy(num_envs, device, seed + 1), 75 * episodes[1] + 1)
    good = bank_io.orient_bank_valid(orient.env)
    out["orient_heaps_per_type"] = int(good.shape[1])
    if save_dir:
        bank_io.save_orient_heap_bank(orient.env, orient.scene, os.path.join(save_dir, "saved_searching_ternimal_states_good_mo_tvalue.pkl"))
    orient.env.close()
    # ---- stage 3: grasp and lift it
    grasp = BlockAssemblyGraspSim(_cfg(num_envs, 150, 1.0), device_id=device_id, seed=seed, heap_bank=good, tvalue_weights=tvalue_weights)
    env = RLgamesVecTaskPython(grasp, device)
    out["grasp_mean_reward"] = _run(env, policies.get("grasp") or _random_policy(num_envs, device, seed + 2), 150 * episodes[2] + 1)
    _, _, idx = grasp.env.grasp_bank()
    torch.cuda.synchronize()
    out["grasp_terminal_states"] = int(idx.sum())
    if save_dir:
        bank_io.save_grasp_bank(grasp.env, os.path.join(save_dir, "saved_grasping_hand_ternimal_states_good_mo_sim.pkl"),
                                os.path.join(save_dir, "saved_grasping_object_ternimal_states_good_mo_sim.pkl"))
    out["banks"] = {"search": heaps, "orient": good}
    grasp.env.close()
    return out
