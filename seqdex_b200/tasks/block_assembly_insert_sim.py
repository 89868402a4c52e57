"""BlockAssemblyInsertSim with the reference's BaseTask surface (BT:24-150; IS = tasks/block_assembly/
allegro_hand_block_assembly_insert_sim.py:96-1725), backed by the CUDA kernels behind the C-ABI (``scene.task = SDX_TASK_INSERT_SIM``,
csrc/sdx_task_insert.cuh).  The last link of the chain (BASELINE ``configs[3]``): every episode starts from a grasp GraspSim banked
(``saved_grasping_{object,hand}_ternimal_states_*.pkl``, IS:372-375, 1449-1453) and has to seat the held brick on the base-plate.

Physics model (DESIGN.md "InsertSim"): FLAT PLATE -- the 4x4x{1,2,4} base-plate (by ``env % 3``, IS:971-977) is a box that ends at
the top of the plate's body; its studs are not collided, so the brick reaches the insertion pose resting on the plate but nothing
snaps or locks sideways.  Observations, reward, resets and reset_idx are the reference's, pinned by goldens.
There is no PyTorch implementation of any phase here."""
from __future__ import annotations

import numpy as np
import torch

from ..env import SdxEnv
from ..randomization import RandomizedTaskMixin
from ..scene import FINGERTIP_BODIES, HAND_BASE_BODY, PREPARE_ARM, FINGER_RESET_UNSCALED, robot_fk
from .cfg import TASK_CFG, scene_from_cfg

DEFAULT_CFG = TASK_CFG["BlockAssemblyInsertSim"]   # cfg/allegro_hand_block_assembly_insert_sim.yaml: env scalars + the whole sim block (tasks/cfg.py)


def synthetic_grasp_bank(scene, per_type=4, seed=0):
    """Stand-in for the unshipped ``saved_grasping_*_ternimal_states_good_mo_sim.pkl`` (IS:372-375) when no GraspSim run hands its
    rings over: the hand at GraspSim's reset pose (GS:1523-1536) with the brick between the fingertips, small pose jitter per row.
    Returns (hand [8, K, 23, 2], obj [8, K, 13])."""
    rng = np.random.default_rng(seed)
    lo, hi = scene.dof_lo, scene.dof_hi
    q = np.concatenate([np.asarray(PREPARE_ARM), 0.5 * (np.asarray(FINGER_RESET_UNSCALED) + 1.0) * (hi[7:] - lo[7:]) + lo[7:]])
    X, R, _, _ = robot_fk(q)
    tips = np.stack([X[b] + R[b] @ np.array([0.0, 0.0, 0.04]) for b in FINGERTIP_BODIES])
    centre = 0.5 * (tips.mean(0) + X[HAND_BASE_BODY])
    hand = np.zeros((8, per_type, 23, 2), np.float32)
    obj = np.zeros((8, per_type, 13), np.float32)
    hand[..., 0] = q.astype(np.float32)
    obj[..., 0:3] = centre.astype(np.float32) + rng.uniform(-0.005, 0.005, size=(8, per_type, 3)).astype(np.float32)
    yaw = rng.uniform(-0.3, 0.3, size=(8, per_type))
    obj[..., 5], obj[..., 6] = np.sin(yaw / 2), np.cos(yaw / 2)
    return hand, obj


class BlockAssemblyInsertSim(RandomizedTaskMixin):
    num_obs_dict = {"partial_contact": 75, "student_partial_contact": 30}      # IS:173-176
    stack_obs = 1                                                              # IS:171

    def __init__(self, cfg=None, sim_params=None, physics_engine=None, device_type="cuda", device_id=0, headless=True,
                 agent_index=None, is_multi_agent=False, grasp_bank=None, bank_per_type=4, seed=22):
        cfg = cfg or DEFAULT_CFG
        self.cfg = cfg
        if device_type not in ("cuda", "GPU"):
            raise RuntimeError("seqdex_b200 runs on CUDA devices only (the reference's --pipeline=cpu has no counterpart here)")
        env_cfg = cfg["env"]
        self.num_envs = int(env_cfg["numEnvs"])
        self.max_episode_length = int(env_cfg.get("episodeLength", DEFAULT_CFG["env"]["episodeLength"]))
        self.control_freq_inv = int(env_cfg.get("controlFrequencyInv", 1))
        self.device = f"cuda:{device_id}"
        self.device_id = device_id
        self.headless = headless
        self.one_frame_num_obs, self.one_frame_num_states = 75, 188
        self.num_obs, self.num_states, self.num_actions = 75, 188, 23               # IS:187-191
        self.scene = scene_from_cfg("BlockAssemblyInsertSim", cfg, seed)
        self.env = SdxEnv(self.scene, self.num_envs, device_id, seed)
        if grasp_bank is None:   # IS:372-375 loads two unshipped pickles; we synthesise the same kind of data
            grasp_bank = synthetic_grasp_bank(self.scene, bank_per_type, seed)
        self.env.set_grasp_bank(*grasp_bank)
        t = self.env.tensor
        self.obs_buf, self.states_buf, self.rew_buf = t("OBS"), t("STATES"), t("REW")
        self.reset_buf, self.progress_buf = t("RESET"), t("PROGRESS")
        self.successes, self.consecutive_successes = t("SUCCESSES"), t("CONSEC")
        self.actions = t("ACTIONS")
        self.segmentation_target_init = t("TARGET_INIT")
        self.success_buf = t("SUCCESS")               # IS:1348-1350: [inserted, not inserted] of the episode that just ended
        self.extra_target_pose = t("PLATE")           # the base-plate's root pose (IS:1438-1446)
        self.rot_err = t("ROT_ERR")
        self.meta_rew_buf = torch.zeros(self.num_envs, device=self.device)
        zeros = torch.zeros(self.num_envs, device=self.device)
        self.extras = {"emergence_reward": zeros, "heap_movement_penalty": zeros, "meta_reward": self.meta_rew_buf,
                       "student_obs_buf": self.obs_buf[:, 0:30], "success_buf": torch.zeros_like(self.reset_buf)}   # IS:456-457, 1070-1072
        self._dr_init(cfg, seed)

    # ---- BaseTask.step (BT:130-150)
    def step(self, actions):
        actions = self._dr_before(actions)
        self.env.step(actions)
        self._dr_after()
        self.meta_rew_buf += self.rew_buf          # IS:1068

    def pre_physics_step(self, actions):
        self.env.pre_physics(actions)

    def post_physics_step(self):
        self.env.post_physics()

    def get_states(self):
        return self.states_buf

    def render(self, sync_frame_time=False):
        return None

    def insert_success_rate(self):
        """mean of success_buf[:, 0] (the reference prints it at every reset, IS:1366)"""
        return float(self.success_buf[:, 0].mean())
