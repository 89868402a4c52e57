"""The task chain Search -> Orient -> GraspSim -> InsertSim (BASELINE ``configs[3]``) with the reference's hand-offs kept on the device
(InsertSim runs on the flat-plate contact model, DESIGN.md "InsertSim").

In the reference each stage is a separate ``train_rlgames.py`` run that ends by pickling what the next one loads:
    Search   writes  intermediate_state/saved_searching_ternimal_states_medium_mo_tvalue.pkl   (SE:1348-1352)
    Orient   reads it (OR:419-420), writes  saved_searching_ternimal_states_good_mo_tvalue.pkl (OR:1510-1512)
    GraspSim reads it (GS:412-413), writes  saved_grasping_{object,hand}_ternimal_states_*.pkl (GS:1448-1451)
    InsertSim reads those (IS:372-375) and restores one banked grasp per episode (IS:1449-1453)
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
from .tasks import BlockAssemblyGraspSim, BlockAssemblyInsertSim, BlockAssemblyOrient, BlockAssemblySearch
from .tasks.block_assembly_insert_sim import synthetic_grasp_bank
from .vec_task import RLgamesVecTaskPython


def _cfg(num_envs, episode, ema):
    return {"env": {"numEnvs": num_envs, "episodeLength": episode, "actionsMovingAverage": ema},
            "sim": {"substeps": 2, "physx": {}}, "task": {"randomize": False}}


def _random_policy(num_envs, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    return lambda obs: torch.rand(num_envs, 23, device=device, generator=g) * 2 - 1


def _run(env, policy, steps, timing=None, name=None):
    """``steps`` VecTask.step calls under ``policy``; with ``timing`` (a dict) the loop is bracketed by CUDA events on the current
    stream and ``timing[name]`` = (seconds, steps) -- environment construction and bank synthesis stay outside"""
    obs = env.reset()
    rew_sum = torch.zeros((), device=obs["obs"].device)
    if timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
    for _ in range(steps):
        obs, rew, _, _ = env.step(policy(obs["obs"]))
        rew_sum += rew.mean()
    if timing is not None:
        e1.record()
        torch.cuda.synchronize()
        timing[name] = (e0.elapsed_time(e1) * 1e-3, steps)
    return float(rew_sum) / max(steps, 1)


def grasp_bank_for_insert(grasp_env, scene, fill=4, seed=0, synthetic=synthetic_grasp_bank):
    """GraspSim's terminal-state rings -> the bank InsertSim restores from: (hand [8, K, 23, 2], obj [8, K, 13]) on the device, K = the
    most grasps any brick type banked; a type that banked fewer cycles through its own rows, a type that banked NONE (an untrained
    policy rarely lifts every type) is filled with the synthetic stand-in grasps, and how many types that were is returned.
    The same hand-over serves ToolPositioningGrasp -> ToolPositioningOrient (``synthetic`` = tasks.tool_positioning.synthetic_tool_grasp_bank)."""
    hand, obj, idx = grasp_env.grasp_bank()
    torch.cuda.synchronize()
    counts = [min(int(c), hand.shape[1]) for c in idx.cpu().tolist()]
    K = max(max(counts), fill)
    sh, so = synthetic(scene, K, seed)
    out_h = torch.from_numpy(sh).to(hand.device)
    out_o = torch.from_numpy(so).to(hand.device)
    for ty, c in enumerate(counts):
        if c > 0:
            sel = torch.arange(K, device=hand.device) % c
            out_h[ty], out_o[ty] = hand[ty, sel], obj[ty, sel]
    return (out_h.contiguous(), out_o.contiguous()), sum(1 for c in counts if c == 0)


def run_chain(num_envs=256, device_id=0, episodes=(2, 2, 1), policies=None, tvalue_weights=None, bank_capacity=64, seed=22,
              save_dir=None, min_bank=1, timing=None):
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
    out["search_mean_reward"] = _run(env, policies.get("search") or _random_policy(num_envs, device, seed), 75 * episodes[0] + 1, timing, "search")
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
    out["orient_mean_reward"] = _run(env, policies.get("orient") or _random_policy(num_envs, device, seed + 1), 75 * episodes[1] + 1, timing, "orient")
    good = bank_io.orient_bank_valid(orient.env)
    out["orient_heaps_per_type"] = int(good.shape[1])
    if save_dir:
        bank_io.save_orient_heap_bank(orient.env, orient.scene, os.path.join(save_dir, "saved_searching_ternimal_states_good_mo_tvalue.pkl"))
    orient.env.close()
    # ---- stage 3: grasp and lift it
    grasp = BlockAssemblyGraspSim(_cfg(num_envs, 150, 1.0), device_id=device_id, seed=seed, heap_bank=good, tvalue_weights=tvalue_weights)
    env = RLgamesVecTaskPython(grasp, device)
    out["grasp_mean_reward"] = _run(env, policies.get("grasp") or _random_policy(num_envs, device, seed + 2), 150 * episodes[2] + 1, timing, "grasp")
    _, _, idx = grasp.env.grasp_bank()
    torch.cuda.synchronize()
    out["grasp_terminal_states"] = int(idx.sum())
    if save_dir:
        bank_io.save_grasp_bank(grasp.env, os.path.join(save_dir, "saved_grasping_hand_ternimal_states_good_mo_sim.pkl"),
                                os.path.join(save_dir, "saved_grasping_object_ternimal_states_good_mo_sim.pkl"))
    out["banks"] = {"search": heaps, "orient": good}
    # ---- stage 4: seat the held brick on the base-plate (every episode starts from a banked grasp)
    gbank, synthetic_types = grasp_bank_for_insert(grasp.env, grasp.scene, seed=seed)
    out["insert_bank_rows_per_type"], out["insert_bank_synthetic_types"] = int(gbank[0].shape[1]), synthetic_types
    grasp.env.close()
    n_ins = episodes[3] if len(episodes) > 3 else episodes[2]
    insert = BlockAssemblyInsertSim(_cfg(num_envs, 125, 1.0), device_id=device_id, seed=seed, grasp_bank=gbank)
    env = RLgamesVecTaskPython(insert, device)
    out["insert_mean_reward"] = _run(env, policies.get("insert") or _random_policy(num_envs, device, seed + 3), 125 * n_ins + 1, timing, "insert")
    torch.cuda.synchronize()
    out["insert_success_rate"] = insert.insert_success_rate()
    out["banks"]["grasp"] = gbank
    insert.env.close()
    return out
