"""BlockAssemblySearch with the reference's BaseTask surface (BT:24-150; SE = tasks/block_assembly/
allegro_hand_block_assembly_search.py:54-1712), backed by the CUDA kernels behind the C-ABI (``scene.task = SDX_TASK_SEARCH``).
BASELINE ``configs[0]`` -- the reference's own CPU-runnable case -- is this task at ``numEnvs: 4``.  The overview camera the
reference renders (SE:873-878) is replaced by ray casting (csrc/sdx_camera.cuh).  There is no PyTorch implementation of any phase
here; the transition-feasibility gate (RetriGraspTValue, 650 -> 1024 -> 512 -> 128 -> 2, TVF:12-28) is evaluated by the
tensor-core MLP on the ``t_value_obs_buf`` the post-physics kernel maintains."""
from __future__ import annotations

import torch

from ..camera import SEARCH_CAMERA, look_at
from ..env import SdxEnv
from ..randomization import RandomizedTaskMixin
from ..scene import Scene
from .cfg import TASK_CFG, scene_from_cfg

DEFAULT_CFG = TASK_CFG["BlockAssemblySearch"]   # cfg/allegro_hand_block_assembly_search.yaml: env scalars + the whole sim block (tasks/cfg.py)


class BlockAssemblySearch(RandomizedTaskMixin):
    num_obs_dict = {"partial_contact": 62}       # SE:149-153
    stack_obs = 3                                # SE:147

    def __init__(self, cfg=None, sim_params=None, physics_engine=None, device_type="cuda", device_id=0, headless=True,
                 agent_index=None, is_multi_agent=False, seed=22, record_heaps=0, tvalue_seed=None):
        cfg = cfg or DEFAULT_CFG
        self.cfg = cfg
        if device_type not in ("cuda", "GPU"):
            raise RuntimeError("seqdex_b200 runs on CUDA devices only (the reference's --pipeline=cpu has no counterpart here)")
        env_cfg, sim_cfg = cfg["env"], cfg.get("sim", {})
        physx = sim_cfg.get("physx", {})
        self.num_envs = int(env_cfg["numEnvs"])
        self.max_episode_length = int(env_cfg.get("episodeLength", 75))
        self.hand_reset_step = int(env_cfg.get("handResetStep", 45))
        self.device = f"cuda:{device_id}"
        self.device_id = device_id
        self.headless = headless
        self.one_frame_num_obs, self.one_frame_num_states = 62, 188
        self.num_obs, self.num_states, self.num_actions = 62 * 3, 188 * 3, 23       # SE:168-175
        self.scene = scene_from_cfg("BlockAssemblySearch", cfg, seed)
        self.env = SdxEnv(self.scene, self.num_envs, device_id, seed)
        self.camera = look_at(**SEARCH_CAMERA)                                       # SE:755-758, 875
        self.env.set_camera(self.camera)
        if record_heaps:     # saved_searching_ternimal_states_list -> ..._medium_mo_tvalue.pkl, read by Orient (SE:1348-1352, OR:419-420)
            self.env.enable_search_bank(record_heaps)
        t = self.env.tensor
        self.obs_buf, self.states_buf, self.rew_buf = t("OBS"), t("STATES"), t("REW")
        self.reset_buf, self.progress_buf = t("RESET"), t("PROGRESS")
        self.successes, self.consecutive_successes = t("SUCCESSES"), t("CONSEC")
        self.actions = t("ACTIONS")
        self.segmentation_target_init = t("TARGET_INIT")
        self.segmentation_features = t("SEG")          # [:, 0] segmentation_object_point_num, [:, 1:3] centre row / column (SE:306-308)
        self.emergence_reward = t("EMERGENCE")
        self.t_value_obs_buf = t("TVOBS")
        self.meta_rew_buf = torch.zeros(self.num_envs, device=self.device)
        self.extras = {"emergence_reward": self.emergence_reward, "meta_reward": self.meta_rew_buf,
                       "student_obs_buf": self.obs_buf[:, 0:30], "success_buf": torch.zeros_like(self.reset_buf)}
        self._gate = None
        self._tvalue_seed = seed if tvalue_seed is None else tvalue_seed
        self._dr_init(cfg, seed)                    # SE:106-107 (task.randomize)

    # ---- BaseTask.step (BT:130-150)
    def step(self, actions):
        actions = self._dr_before(actions)         # BT:131-132 (only with task.randomize)
        self.env.step(actions)
        self._dr_after()                          # BT:149-150
        self.meta_rew_buf += self.rew_buf          # SE:954

    def pre_physics_step(self, actions):
        self.env.pre_physics(actions)

    def post_physics_step(self):
        self.env.post_physics()

    @property
    def tvalue(self):
        """sigmoid(RetriGraspTValue(t_value_obs_buf))[:, 1] (SE:1135-1136); ELU on the output layer as in TVF:26"""
        from ..ppo import MLP
        if self._gate is None:
            self._gate = MLP(650, 2, max(128, (self.num_envs + 7) // 8 * 8), device=self.device_id, seed=self._tvalue_seed, hidden=(1024, 512, 128))
        z = self._gate.forward(self.t_value_obs_buf)
        return torch.sigmoid(torch.nn.functional.elu(z))[:, 1]

    def get_states(self):
        return self.states_buf

    def render(self, sync_frame_time=False):
        return None
