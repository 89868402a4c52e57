"""BlockAssemblyGraspSim with the reference's BaseTask surface (BT:24-150, GS:96-1776), backed by the
CUDA kernels behind the C-ABI.  Same constructor signature, same buffers (obs_buf, states_buf, rew_buf,
reset_buf, progress_buf, extras), same step() phase order -- so ``RLgamesVecTaskPython`` and an
rl_games-style runner drive it unchanged.  There is no PyTorch implementation of any phase here."""
from __future__ import annotations

import torch

from ..env import SdxEnv, make_heap_bank
from ..randomization import RandomizedTaskMixin
from ..scene import Scene
from .cfg import TASK_CFG, scene_from_cfg


DEFAULT_CFG = TASK_CFG["BlockAssemblyGraspSim"]   # cfg/allegro_hand_block_assembly_grasp_sim.yaml: env scalars + the whole sim block (tasks/cfg.py)


class BlockAssemblyGraspSim(RandomizedTaskMixin):
    num_obs_dict = {"partial_contact": 132}      # GS:191-195
    stack_obs = 3                                # GS:189

    def __init__(self, cfg=None, sim_params=None, physics_engine=None, device_type="cuda", device_id=0, headless=True,
                 agent_index=None, is_multi_agent=False, heap_bank=None, bank_per_type=64, seed=22, tvalue_weights=None):
        cfg = cfg or DEFAULT_CFG
        self.cfg = cfg
        if device_type not in ("cuda", "GPU"):
            raise RuntimeError("seqdex_b200 runs on CUDA devices only (the reference's --pipeline=cpu has no counterpart here)")
        env_cfg, sim_cfg = cfg["env"], cfg.get("sim", {})
        physx = sim_cfg.get("physx", {})
        self.num_envs = int(env_cfg["numEnvs"])
        self.max_episode_length = int(env_cfg.get("episodeLength", 150))
        self.control_freq_inv = int(env_cfg.get("controlFrequencyInv", 1))
        self.device = f"cuda:{device_id}"
        self.device_id = device_id
        self.headless = headless
        self.one_frame_num_obs, self.one_frame_num_states = 132, 188
        self.num_obs, self.num_states, self.num_actions = 132 * 3, 188 * 3, 23      # GS:209-211
        self.scene = scene_from_cfg("BlockAssemblyGraspSim", cfg, seed)
        self.env = SdxEnv(self.scene, self.num_envs, device_id, seed)
        if heap_bank is None:   # GS:412-413 loads an unshipped pickle; we synthesise the same kind of data
            heap_bank = make_heap_bank(self.scene, bank_per_type, device_id, seed=seed)
        self.env.set_heap_bank(heap_bank)
        if tvalue_weights is None:
            tvalue_weights = default_tvalue_weights(seed)
        self.env.set_tvalue_weights(tvalue_weights)
        t = self.env.tensor
        self.obs_buf, self.states_buf, self.rew_buf = t("OBS"), t("STATES"), t("REW")           # BT:57-62
        self.reset_buf, self.progress_buf = t("RESET"), t("PROGRESS")                            # BT:63-66
        self.successes, self.consecutive_successes, self.tvalue = t("SUCCESSES"), t("CONSEC"), t("TVALUE")
        self.actions = t("ACTIONS")
        self.segmentation_target_init = t("TARGET_INIT")
        self.meta_rew_buf = torch.zeros(self.num_envs, device=self.device)
        zeros = torch.zeros(self.num_envs, device=self.device)
        self.extras = {"emergence_reward": zeros, "heap_movement_penalty": zeros, "meta_reward": self.meta_rew_buf,
                       "student_obs_buf": self.obs_buf[:, 0:30], "success_buf": torch.zeros_like(self.reset_buf)}   # GS:458-459, 1071-1073
        self._dr_init(cfg, seed)                    # GS:106-107 / OR:106-107 (task.randomize)

    # ---- BaseTask.step (BT:130-150)
    def step(self, actions):
        actions = self._dr_before(actions)         # BT:131-132 (only with task.randomize)
        self.env.step(actions)
        self._dr_after()                          # BT:149-150
        self.meta_rew_buf += self.rew_buf          # GS:1069

    def pre_physics_step(self, actions):
        self.env.pre_physics(actions)

    def post_physics_step(self):
        self.env.post_physics()

    def get_states(self):
        return self.states_buf

    def render(self, sync_frame_time=False):
        return None


def default_tvalue_weights(seed=0):
    """torch-default initialisation of GraspInsertTValue(4, 2) (TVF:30-46), flat state_dict order."""
    import math
    g = torch.Generator().manual_seed(seed)
    out = []
    for (o, i) in ((256, 4), (128, 256), (64, 128), (2, 64)):
        bound = 1.0 / math.sqrt(i)
        out += [((torch.rand(o, i, generator=g) * 2 - 1) * bound).reshape(-1), (torch.rand(o, generator=g) * 2 - 1) * bound]
    return torch.cat(out).numpy()
