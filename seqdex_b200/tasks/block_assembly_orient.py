"""BlockAssemblyOrient with the reference's BaseTask surface (BT:24-150; OR = tasks/block_assembly/
allegro_hand_block_assembly_orient.py:94-1934), backed by the CUDA kernels behind the C-ABI (``scene.task = SDX_TASK_ORIENT``).
Same scene and contact step as GraspSim; its own finger gains (OR:588-598), action mapping, 62 x 3 observations, reward and the
scripted reset (csrc/sdx_task_orient.cuh).  There is no PyTorch implementation of any phase here."""
from __future__ import annotations

import torch

from ..env import SdxEnv, make_heap_bank
from ..randomization import RandomizedTaskMixin
from ..scene import Scene
from .cfg import TASK_CFG, scene_from_cfg
from .block_assembly_grasp_sim import default_tvalue_weights

DEFAULT_CFG = TASK_CFG["BlockAssemblyOrient"]   # cfg/allegro_hand_block_assembly_orient.yaml: env scalars + the whole sim block (tasks/cfg.py)


class BlockAssemblyOrient(RandomizedTaskMixin):
    num_obs_dict = {"partial_contact": 62, "student_partial_contact": 30}      # OR:189-192
    stack_obs = 3                                                              # OR:187

    def __init__(self, cfg=None, sim_params=None, physics_engine=None, device_type="cuda", device_id=0, headless=True,
                 agent_index=None, is_multi_agent=False, heap_bank=None, bank_per_type=64, seed=22, tvalue_weights=None,
                 record_heaps=0):
        cfg = cfg or DEFAULT_CFG
        self.cfg = cfg
        if device_type not in ("cuda", "GPU"):
            raise RuntimeError("seqdex_b200 runs on CUDA devices only (the reference's --pipeline=cpu has no counterpart here)")
        env_cfg, sim_cfg = cfg["env"], cfg.get("sim", {})
        physx = sim_cfg.get("physx", {})
        self.num_envs = int(env_cfg["numEnvs"])
        self.max_episode_length = int(env_cfg.get("episodeLength", 75))
        self.control_freq_inv = int(env_cfg.get("controlFrequencyInv", 1))
        self.device = f"cuda:{device_id}"
        self.device_id = device_id
        self.headless = headless
        self.one_frame_num_obs, self.one_frame_num_states = 62, 188
        self.num_obs, self.num_states, self.num_actions = 62 * 3, 188 * 3, 23       # OR:206-208
        self.scene = scene_from_cfg("BlockAssemblyOrient", cfg, seed)
        self.env = SdxEnv(self.scene, self.num_envs, device_id, seed)
        if heap_bank is None:   # OR:419-420 loads the pickle Search writes (unshipped); we synthesise the same kind of data
            heap_bank = make_heap_bank(self.scene, bank_per_type, device_id, seed=seed)
        self.env.set_heap_bank(heap_bank)
        self.env.set_tvalue_weights(default_tvalue_weights(seed) if tvalue_weights is None else tvalue_weights)
        if record_heaps:        # saved_digging_ternimal_states_list -> saved_searching_ternimal_states_good_mo_tvalue.pkl (OR:1465-1513)
            self.env.enable_orient_heap_bank(record_heaps)
        t = self.env.tensor
        self.obs_buf, self.states_buf, self.rew_buf = t("OBS"), t("STATES"), t("REW")
        self.reset_buf, self.progress_buf = t("RESET"), t("PROGRESS")
        self.successes, self.consecutive_successes, self.tvalue = t("SUCCESSES"), t("CONSEC"), t("TVALUE")
        self.actions = t("ACTIONS")
        self.segmentation_target_init = t("TARGET_INIT")
        self.meta_rew_buf = torch.zeros(self.num_envs, device=self.device)
        zeros = torch.zeros(self.num_envs, device=self.device)
        self.extras = {"emergence_reward": zeros, "heap_movement_penalty": zeros, "meta_reward": self.meta_rew_buf,
                       "student_obs_buf": self.obs_buf[:, 0:30], "success_buf": torch.zeros_like(self.reset_buf)}   # OR:478-479, 1070-1072
        self._dr_init(cfg, seed)                    # GS:106-107 / OR:106-107 (task.randomize)

    # ---- BaseTask.step (BT:130-150)
    def step(self, actions):
        actions = self._dr_before(actions)         # BT:131-132 (only with task.randomize)
        self.env.step(actions)
        self._dr_after()                          # BT:149-150
        self.meta_rew_buf += self.rew_buf          # OR:1068

    def pre_physics_step(self, actions):
        self.env.pre_physics(actions)

    def post_physics_step(self):
        self.env.post_physics()

    def get_states(self):
        return self.states_buf

    def render(self, sync_frame_time=False):
        return None
