"""The bi-directional optimisation schedule of ``scripts/bi_optimization.py`` (BO:36-134) as a function over this engine.

The reference's script loops ten times over
    forward initialisation : train Search -> Orient -> GraspSim -> InsertSim, each stage started from what the one before it banked
    backward fine-tuning   : InsertSim with its transition-feasibility (t-value) function, then ``TValue_Trainer`` on the rows it
                             recorded; GraspSim likewise; Orient likewise (BO:121-126)
by calling ``main_rlgames(task, num_envs, use_t_value, policy_path)`` (BO:36-108: build the task, run rl_games' Runner, return the
checkpoint path) and ``transition_value_trainer(task, rollout)`` (BO:110-113).  As written it cannot complete -- Search and Orient
``exit()`` from inside ``reset_idx`` once their banks are full (SE:1355, OR:1515) -- so, as the survey says,
it is treated as the SPECIFICATION of the schedule: the same two functions, the same order, here with the banks handed over on the
device (``chain.py``) and the rows the t-value trainer needs recorded by the env (``sdx_tvalue_dataset``) or gathered from the task's
buffers instead of going through HDF5 files.

``ToolPositioning`` (BO:127-134) is the same loop over two tasks: ToolPositioningGrasp and ToolPositioningOrient forward (Orient starts
from the grasps Grasp banked), Orient backward with its rows recorded, then the t-value fit.  What Orient records (TO:1287-1296) is the
tool's pose at the START of the episode, 7 wide, labelled by the episode's success; ``TValue_Trainer`` holds 4-wide rows (TVT:158-159) and
re-normalises each row to a unit quaternion (TVT:217), so as written the 7-wide assignment fails -- here the quaternion part of the
pose is what is fitted.

Nothing here computes: PPO is ``ppo.A2CAgent`` (CUDA kernels), the t-value fit is ``tvalue.TValueTrainer``.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import bank_io
from .chain import grasp_bank_for_insert
from .ppo import A2CAgent, PPOConfig
from .tasks import (BlockAssemblyGraspSim, BlockAssemblyInsertSim, BlockAssemblyOrient, BlockAssemblySearch, ToolPositioningGrasp,
                    ToolPositioningOrient)
from .tasks.tool_positioning import synthetic_tool_grasp_bank
from .tvalue import TValueTrainer
from .vec_task import RLgamesVecTaskPython

STAGES = ("BlockAssemblySearch", "BlockAssemblyOrient", "BlockAssemblyGraspSim", "BlockAssemblyInsertSim")
TOOL_STAGES = ("ToolPositioningGrasp", "ToolPositioningOrient")
# the PPO yaml main_rlgames picks per task (BO:44-52): ppo_continuous_insert.yaml = minibatch 4096, critic_coef 4 -- the value loss is
# the central-value net's here, which has its own optimiser, so only the minibatch differs
_PPO = {"BlockAssemblyInsertSim": dict(minibatch_size=4096)}


class StageState:
    """what one stage leaves for the next (the reference's pickles under intermediate_state/), resident on the device"""

    def __init__(self):
        self.heaps_medium = None      # Search  -> Orient   : saved_searching_ternimal_states_medium_mo_tvalue.pkl
        self.heaps_good = None        # Orient  -> GraspSim : saved_searching_ternimal_states_good_mo_tvalue.pkl
        self.grasps = None            # GraspSim -> InsertSim: saved_grasping_{hand,object}_ternimal_states_good_mo_sim.pkl
        self.tool_grasps = None       # ToolPositioningGrasp -> Orient: saved_orient_grasp_{object,hand}_init_tvalue_temporal.pkl (TO:365-368)
        self.tvalue = {}              # task -> flat GraspInsertTValue weights fitted by transition_value_trainer
        self.datasets = {}            # task -> (success rows [n, 4], failure rows [m, 4])


def _cfg(num_envs):
    return {"env": {"numEnvs": num_envs}, "sim": {"physx": {}}, "task": {"randomize": False}}


def _build(task, num_envs, state, device_id, seed, bank_capacity):
    if task == "BlockAssemblySearch":
        return BlockAssemblySearch(_cfg(num_envs), device_id=device_id, seed=seed, record_heaps=bank_capacity)
    if task == "BlockAssemblyOrient":
        return BlockAssemblyOrient(_cfg(num_envs), device_id=device_id, seed=seed, heap_bank=state.heaps_medium, record_heaps=bank_capacity,
                                   tvalue_weights=state.tvalue.get("BlockAssemblyGraspSim"))          # its gate asks: can GraspSim start from here?
    if task == "BlockAssemblyGraspSim":
        t = BlockAssemblyGraspSim(_cfg(num_envs), device_id=device_id, seed=seed, heap_bank=state.heaps_good,
                                  tvalue_weights=state.tvalue.get("BlockAssemblyInsertSim"))           # ... can InsertSim start from this grasp?
        t.env.enable_tvalue_dataset(65536)                                                             # GS:1402-1438 (save_hdf5)
        return t
    if task == "BlockAssemblyInsertSim":
        return BlockAssemblyInsertSim(_cfg(num_envs), device_id=device_id, seed=seed, grasp_bank=state.grasps)
    if task == "ToolPositioningGrasp":
        return ToolPositioningGrasp(_cfg(num_envs), device_id=device_id, seed=seed)
    if task == "ToolPositioningOrient":
        return ToolPositioningOrient(_cfg(num_envs), device_id=device_id, seed=seed, grasp_bank=state.tool_grasps)
    raise ValueError(task)


def _harvest(task_name, task, state, seed):
    """what the reference's reset_idx pickles / writes to HDF5 when a run ends"""
    env = task.env
    if task_name == "BlockAssemblySearch":
        try:
            state.heaps_medium = bank_io.search_bank_valid(env)
        except RuntimeError:                  # too short a run for every brick type to have been dug out once: settled synthetic heaps stand in
            from .env import make_heap_bank
            state.heaps_medium = make_heap_bank(task.scene, 4, env.device_index, seed=seed)
    elif task_name == "BlockAssemblyOrient":
        try:
            state.heaps_good = bank_io.orient_bank_valid(env)
        except RuntimeError:                  # an untrained policy may leave a brick type without a face-up heap: GraspSim then starts from Search's
            state.heaps_good = state.heaps_medium
    elif task_name == "BlockAssemblyGraspSim":
        state.grasps, _ = grasp_bank_for_insert(env, task.scene, seed=seed)
        s, f, _ = env.tvalue_dataset()
        state.datasets[task_name] = (s.clone(), f.clone())
    elif task_name == "ToolPositioningGrasp":
        state.tool_grasps, _ = grasp_bank_for_insert(env, task.scene, seed=seed, synthetic=synthetic_tool_grasp_bank)


def main_rlgames(task, num_envs, use_t_value=False, policy_path="", state=None, iterations=4, device_id=0, seed=22, work_dir="runs",
                 bank_capacity=64):
    """BO:36-108.  Builds ``task`` on ``num_envs`` envs from what the previous stages left in ``state``, restores ``policy_path`` if
    given ("Base" = none, BO:42-43), trains ``iterations`` PPO iterations and saves ``<work_dir>/<task>/nn/<task>.pth``.  Returns the
    checkpoint path (the reference returns it only when ``use_t_value`` is False; the caller here ignores it in that case too).
    With ``use_t_value`` the stage's gate uses the fitted t-value function of the NEXT stage (``state.tvalue``) and, for InsertSim, the
    rows a later ``transition_value_trainer`` needs are gathered: the camera-frame target quaternion at the end of every episode,
    labelled by ``success_buf`` (IS:1388-1403)."""
    state = state or StageState()
    t = _build(task, num_envs, state, device_id, seed, bank_capacity)
    venv = RLgamesVecTaskPython(t, t.device)
    agent = A2CAgent(venv, PPOConfig(minibatch_size=min(_PPO.get(task, {}).get("minibatch_size", 8 * num_envs), 8 * num_envs), seed=seed),
                     device=device_id)
    if policy_path and policy_path != "Base":
        agent.restore(policy_path)
    rows_s, rows_f = [], []
    info = {}
    steps_done = 0
    record = use_t_value and task in ("BlockAssemblyInsertSim", "ToolPositioningOrient")
    for _ in range(iterations):
        if record:
            # one rollout step at a time so that the rows can be gathered where the reference gathers them (in reset_idx)
            if agent.obs is None:
                first = venv.reset()
                agent.set_obs(first["obs"], first["states"])
            for k in range(agent.H):
                a = agent.act(k)
                if task == "ToolPositioningOrient":
                    qcam = t.segmentation_target_init[:, 3:7].clone()        # t_value_obs_buf: the pose the episode STARTED from (TO:1290, 1400)
                else:
                    qcam = t.states_buf[:, 177:181].clone()                  # camera_view_segmentation_target_rot of the state the episode may end in
                o, rew, dones, _ = venv.step(a)
                agent.next_obs.copy_(o["obs"]); agent.next_states.copy_(o["states"])
                agent.record(k, rew, dones)
                ended = t.progress_buf == 1                                   # envs whose reset_idx ran in this step (progress 0 -> 1): success_buf is theirs (IS:1348-1350)
                steps_done += 1
                if steps_done > 1 and bool(ended.any()):                      # the resets of the very first step end no episode (total_steps > 0, IS:1388, TO:1287)
                    ok = t.success_buf[:, 0] > 0.5
                    rows_s.append(qcam[ended & ok]); rows_f.append(qcam[ended & ~ok])
            agent.finish_rollout()
            info = agent.update()
        else:
            info = agent.train_epoch()
    if rows_s:
        state.datasets[task] = (torch.cat(rows_s), torch.cat(rows_f))
    _harvest(task, t, state, seed)
    path = os.path.join(work_dir, task, "nn", task)
    path = agent.save(path)
    t.env.close()
    return path, info


def transition_value_trainer(task, rollout, state, device_id=0, seed=22, min_rows=64):
    """BO:110-113 -> TVT:127-248: fit GraspInsertTValue on the rows ``task`` recorded; the weights go to ``state.tvalue[task]`` (the
    reference saves ``grasp_insert_TValue_*.pt`` under intermediate_state/<task>_t_value/).  Returns the validation accuracy, or None
    when the stage recorded too few rows of either label to train on (an untrained policy rarely succeeds)."""
    s, f = state.datasets.get(task, (None, None))
    if s is None or len(s) < min_rows or len(f) < min_rows:
        return None
    tr = TValueTrainer(s.detach().cpu().numpy(), f.detach().cpu().numpy(), device=device_id, seed=seed)
    tr.train_rollout(rollout)
    state.tvalue[task] = tr.weights()
    return tr.validate()


def bi_optimization(tasks="BlockAssembly", rounds=10, num_envs=None, iterations=4, tvalue_rollout=10000, device_id=0, seed=22, work_dir="runs",
                    log=None):
    """BO:115-134.  ``num_envs``: dict task -> envs (the script's 128 / 512 / 512 / 512 by default)."""
    if tasks not in ("BlockAssembly", "ToolPositioning"):
        raise Exception("Unrecognized task!")                                 # BO:136-138
    ne = {"BlockAssemblySearch": 128, "BlockAssemblyOrient": 512, "BlockAssemblyGraspSim": 512, "BlockAssemblyInsertSim": 512,
          "ToolPositioningGrasp": 512, "ToolPositioningOrient": 512}
    ne.update(num_envs or {})
    state = StageState()
    history = []
    say = log or (lambda *a: None)
    kw = dict(iterations=iterations, device_id=device_id, seed=seed, work_dir=work_dir)
    if tasks == "ToolPositioning":                                            # BO:127-134
        for i in range(rounds):
            rec = {"round": i}
            paths = {}
            for task in TOOL_STAGES:
                paths[task], info = main_rlgames(task, ne[task], state=state, **kw)
                rec[f"forward/{task}"] = info.get("mean_reward")
                say(i, "forward", task, info)
            task = "ToolPositioningOrient"
            _, info = main_rlgames(task, ne[task], use_t_value=True, policy_path=paths[task], state=state, **kw)
            rec[f"backward/{task}"] = info.get("mean_reward")
            rec[f"tvalue/{task}"] = transition_value_trainer(task, tvalue_rollout, state, device_id, seed)
            say(i, "backward", task, info, rec[f"tvalue/{task}"])
            history.append(rec)
        return history, state
    for i in range(rounds):
        rec = {"round": i}
        # forward initialisation (BO:118-121)
        paths = {}
        for task in STAGES:
            paths[task], info = main_rlgames(task, ne[task], state=state, **kw)
            rec[f"forward/{task}"] = info.get("mean_reward")
            say(i, "forward", task, info)
        # backward fine-tuning (BO:123-128; Orient at 128 envs, BO:127)
        for task, n in (("BlockAssemblyInsertSim", ne["BlockAssemblyInsertSim"]), ("BlockAssemblyGraspSim", ne["BlockAssemblyGraspSim"]),
                        ("BlockAssemblyOrient", 128)):
            _, info = main_rlgames(task, n, use_t_value=True, policy_path=paths[task], state=state, **kw)
            rec[f"backward/{task}"] = info.get("mean_reward")
            rec[f"tvalue/{task}"] = transition_value_trainer(task, tvalue_rollout, state, device_id, seed)
            say(i, "backward", task, info, rec[f"tvalue/{task}"])
        history.append(rec)
    return history, state
