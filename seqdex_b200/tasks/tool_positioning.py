"""ToolPositioningGrasp / ToolPositioningOrient (BASELINE ``configs[4]``) with the reference's BaseTask surface (BT:24-150;
TG = tasks/tool_positioning/allegro_hand_tool_positioning_grasp.py:130-1919, TO = tasks/tool_positioning/
allegro_hand_tool_positioning_orient.py:77-1652), backed by the CUDA kernels behind the C-ABI (``scene.task = SDX_TASK_TOOL_GRASP /
SDX_TASK_TOOL_ORIENT``, csrc/sdx_task_tool.cuh).

Grasp picks the hammer up from the bin and banks the good grasps (lifted, fingers on it, within 1 rad of the plate's orientation) into
per-type rings -- what the reference pickles as ``saved_orient_grasp_{object,hand}_init_tvalue_temporal.pkl`` (TG:1459-1467); Orient
starts every episode from one of those grasps (TO:365-368, 1393-1398) and has to turn the hammer in the hand until it is aligned with
the plate.

Physics model (DESIGN.md "ToolPositioning"): the hammer is a COMPOUND of two boxes (handle + head, scene.HAMMER_BOXES) with the URDF's
density; Isaac Gym collides a V-HACD decomposition of harmmer.obj.  Observations, rewards, resets, banking and both reset_idx are the
reference's, pinned by goldens (tests/golden/tool_*.npz).  There is no PyTorch implementation of any phase here."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib
from ..env import SdxEnv
from ..randomization import RandomizedTaskMixin
from ..scene import FINGERTIP_BODIES, HAND_BASE_BODY, TOOL_DEFAULT_ARM, FINGER_RESET_UNSCALED, robot_fk, quat_from_euler_zyx
from .cfg import TASK_CFG, scene_from_cfg


def synthetic_tool_grasp_bank(scene, per_type=4, seed=0):
    """Stand-in for the unshipped ``saved_orient_grasp_{object,hand}_init_tvalue_temporal.pkl`` (TO:365-368) when no ToolPositioningGrasp
    run hands its rings over: the hand at Grasp's reset pose (TG:1536-1548) with the hammer's handle between the fingertips, roughly in
    the plate's orientation, small jitter per row, at rest.  Returns (hand [8, K, 23, 2], obj [8, K, 13])."""
    rng = np.random.default_rng(seed)
    lo, hi = scene.dof_lo, scene.dof_hi
    q = np.concatenate([np.asarray(TOOL_DEFAULT_ARM), 0.5 * (np.asarray(FINGER_RESET_UNSCALED) + 1.0) * (hi[7:] - lo[7:]) + lo[7:]])
    X, R, _, _ = robot_fk(q)
    tips = np.stack([X[b] + R[b] @ np.array([0.0, 0.0, 0.04]) for b in FINGERTIP_BODIES])
    centre = 0.5 * (tips.mean(0) + X[HAND_BASE_BODY])
    hand = np.zeros((8, per_type, 23, 2), np.float32)
    obj = np.zeros((8, per_type, 13), np.float32)
    hand[..., 0] = q.astype(np.float32)
    obj[..., 0:3] = centre.astype(np.float32) + rng.uniform(-0.005, 0.005, size=(8, per_type, 3)).astype(np.float32)
    base = np.asarray(quat_from_euler_zyx(0.0, 3.1415, 0.0))
    for t in range(8):
        for k in range(per_type):
            a = rng.uniform(-0.4, 0.4)
            dq = np.array([0.0, 0.0, np.sin(a / 2), np.cos(a / 2)])
            x1, y1, z1, w1 = base; x2, y2, z2, w2 = dq
            obj[t, k, 3:7] = [w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                              w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2]
    return hand, obj


class _ToolPositioning(RandomizedTaskMixin):
    num_obs_dict = {"partial_contact": 156, "student_partial_contact": 30}     # TG:224-227, TO:170-173
    stack_obs = 3                                                              # TG:222, TO:168
    TASK = None

    def _setup(self, cfg, device_type, device_id, headless, seed):
        cfg = cfg or TASK_CFG[self.TASK]
        self.cfg = cfg
        if device_type not in ("cuda", "GPU"):
            raise RuntimeError("seqdex_b200 runs on CUDA devices only (the reference's --pipeline=cpu has no counterpart here)")
        env_cfg = cfg["env"]
        self.num_envs = int(env_cfg["numEnvs"])
        self.max_episode_length = int(env_cfg.get("episodeLength", TASK_CFG[self.TASK]["env"]["episodeLength"]))
        self.control_freq_inv = int(env_cfg.get("controlFrequencyInv", 1))
        self.device = f"cuda:{device_id}"
        self.device_id = device_id
        self.headless = headless
        self.one_frame_num_obs, self.one_frame_num_states = 156, 188
        self.num_obs, self.num_states, self.num_actions = 156 * 3, 188 * 3, 23      # TG:238-243
        self.scene = scene_from_cfg(self.TASK, cfg, seed)
        self.env = SdxEnv(self.scene, self.num_envs, device_id, seed)

    def _bind(self, cfg, seed):
        t = self.env.tensor
        self.obs_buf, self.states_buf, self.rew_buf = t("OBS"), t("STATES"), t("REW")
        self.reset_buf, self.progress_buf = t("RESET"), t("PROGRESS")
        self.successes, self.consecutive_successes = t("SUCCESSES"), t("CONSEC")
        self.actions = t("ACTIONS")
        self.segmentation_target_init = t("TARGET_INIT")
        self.success_buf = t("SUCCESS")               # column 0: the episode that just ended finished aligned with the plate (TG:1428, TO:1282)
        self.extra_target_pose = t("PLATE")           # the plate's root pose (TG:1505-1512)
        self.meta_rew_buf = torch.zeros(self.num_envs, device=self.device)
        zeros = torch.zeros(self.num_envs, device=self.device)
        self.extras = {"emergence_reward": zeros, "heap_movement_penalty": zeros, "meta_reward": self.meta_rew_buf,
                       "student_obs_buf": self.obs_buf[:, 0:30], "success_buf": torch.zeros_like(self.reset_buf)}
        self._dr_init(cfg or TASK_CFG[self.TASK], seed)

    # ---- BaseTask.step (BT:130-150)
    def step(self, actions):
        actions = self._dr_before(actions)
        self.env.step(actions)
        self._dr_after()
        self.meta_rew_buf += self.rew_buf          # TG:1087

    def pre_physics_step(self, actions):
        self.env.pre_physics(actions)

    def post_physics_step(self):
        self.env.post_physics()

    def get_states(self):
        return self.states_buf

    def render(self, sync_frame_time=False):
        return None

    def success_rate(self):
        """mean of success_buf[:, 0] (the reference prints it at every reset, TG:1429)"""
        return float(self.success_buf[:, 0].mean())


class ToolPositioningGrasp(_ToolPositioning):
    TASK = "ToolPositioningGrasp"

    def __init__(self, cfg=None, sim_params=None, physics_engine=None, device_type="cuda", device_id=0, headless=True,
                 agent_index=None, is_multi_agent=False, seed=22):
        self._setup(cfg, device_type, device_id, headless, seed)
        self._bind(cfg, seed)

    def grasp_bank(self):
        """the rings reset_idx fills (TG:1436-1457): (hand [8, 11024, 23, 2], obj [8, 11024, 13], index [8]) on the device"""
        return self.env.grasp_bank()


class ToolPositioningChain(ToolPositioningGrasp):
    """The inner-policy call of ``ToolPositioningChain`` (TC = tasks/tool_positioning/allegro_hand_tool_positioning_chain.py:1733-1768) on
    ToolPositioningGrasp's scene, observations, reward and reset: when env 0's clock reads 118, ``pre_physics_step`` runs 125 steps of a
    FROZEN ToolPositioningOrient policy inside the outer step -- ``insert_policy.predict`` on the clamped ``insertion_obs_buf``
    (TC:1742; the buffer is NOT refreshed inside the loop, so the policy sees one observation 125 times and only its action noise
    varies -- reproduced as written), fingers = the scaled actions, arm = its previous target, contact step -- and
    ``compute_observations`` additionally fills ``insertion_obs_buf`` (TC:1306, 1404-1440).

    What TC changes beyond that is NOT built: its three tool types with fixed bases (TC:800, 816), the ``replan`` bookkeeping and its
    ``success_buf`` variant (TC:1488-1518), the bank gate without the orientation test (TC:1531-1536) and the mid-air teleport of
    resetting tools (TC:1813-1825) -- evaluation scaffolding around checkpoints that live in the authors' home directory (TC:558)."""
    TASK = "ToolPositioningGrasp"
    INNER_AT, INNER_STEPS, INNER_EPISODE = 118, 125, 125               # TC:1733-1734, 561

    def __init__(self, cfg=None, sim_params=None, physics_engine=None, device_type="cuda", device_id=0, headless=True,
                 agent_index=None, is_multi_agent=False, seed=22, insert_policy=None, insert_policy_path=None):
        super().__init__(cfg, sim_params, physics_engine, device_type, device_id, headless, agent_index, is_multi_agent, seed)
        from ..policy_sequencing import NNController
        self.insertion_num_obs = 3 * 156                                 # TC:533-537
        self.insertion_obs_buf = torch.zeros(self.num_envs, self.insertion_num_obs, device=self.device)
        self.insertion_actions = torch.zeros(self.num_envs, 23, device=self.device)
        self.insertion_progress_buf = torch.zeros(self.num_envs, dtype=torch.int64, device=self.device)
        # insert_network.yaml: units [1024, 512, 256] (utils/robot_controller/nn_controller.py:7-58)
        self.insert_policy = insert_policy or NNController(num_actors=self.num_envs, units=(1024, 512, 256), obs_dim=self.insertion_num_obs,
                                                           device=device_id, seed=seed)
        if insert_policy_path:
            self.insert_policy.load(insert_policy_path)                  # TC:558
        self.inner_calls = 0

    def pre_physics_step(self, actions):
        inner = False
        if int(self.progress_buf[0]) == self.INNER_AT and not bool(self.reset_buf[0]):   # TC:1733 reads env 0's clock on the host (after reset_idx)
            inner = True
            dof = self.env.tensor("DOF")
            prev_arm, was_reset = dof[:, 2, :7].clone(), self.reset_buf.bool().clone()
        self.env.pre_physics(actions)                                    # reset_idx, then TG's targets (TC:1688-1731)
        if inner:
            # inside the loop the arm holds prev_targets (TC:1752-1754), which at this point still are the PREVIOUS step's targets
            # (they are only updated at TC:1776) -- for an env reset_idx has just reset, the pose it was reset to
            dof[:, 2, :7] = torch.where(was_reset[:, None], dof[:, 0, :7], prev_arm)
            for _ in range(self.INNER_STEPS):
                self.insertion_progress_buf.zero_()                      # TC:1739 (inside the loop, as written)
                self.insertion_actions = self.insert_policy.predict(torch.clamp(self.insertion_obs_buf, -5.0, 5.0), deterministic=False).contiguous()
                self.env.tool_inner_step(self.insertion_actions)
                self.insertion_progress_buf += 1                         # TC:1768
            self.inner_calls += 1

    def post_physics_step(self):
        self.env.post_physics()
        self.env.tool_insertion_obs(self.insertion_actions, self.insertion_progress_buf, self.insertion_obs_buf, self.INNER_EPISODE)   # TC:1306

    def step(self, actions):
        actions = self._dr_before(actions)
        self.pre_physics_step(actions)
        self.env.simulate()
        self.post_physics_step()
        self._dr_after()
        self.meta_rew_buf += self.rew_buf


class ToolPositioningOrient(_ToolPositioning):
    TASK = "ToolPositioningOrient"

    def __init__(self, cfg=None, sim_params=None, physics_engine=None, device_type="cuda", device_id=0, headless=True,
                 agent_index=None, is_multi_agent=False, grasp_bank=None, bank_per_type=4, seed=22, if_t_value=False):
        self._setup(cfg, device_type, device_id, headless, seed)
        if grasp_bank is None:   # TO:365-368 loads two unshipped pickles; we synthesise the same kind of data
            grasp_bank = synthetic_tool_grasp_bank(self.scene, bank_per_type, seed)
        self.env.set_grasp_bank(*grasp_bank)
        self._bind(cfg, seed)
        # the online t-value update inside reset_idx (TO:1305-1350).  The reference hard-wires `self.if_t_value = False` (TO:377), so it
        # is opt-in here: GraspInsertTValue(input_dim=7) on the tensor-core MLP, Adam(lr=3e-4), BCE-with-logits (TO:379-387)
        self.if_t_value = bool(if_t_value)
        self.total_steps = 0
        if self.if_t_value:
            from ..ppo import MLP
            self.t_value = MLP(7, 2, self.num_envs, device=device_id, seed=seed, hidden=(256, 128, 64))
            self._tv_label = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
            self._tv_dz = torch.zeros(self.num_envs, 2, device=self.device)
            self._tv_stats = torch.zeros(4, device=self.device)

    def online_t_value_update(self, steps=5, lr=0.0003):
        """TO:1305-1350: labels for ALL envs from their current state (``sdx_tool_tvalue_labels``: within 1 cm and 0.1 rad of the plate's
        pose or its pi-about-z twin; also what ``success_buf`` holds afterwards), rows = ``t_value_obs_buf`` = the pose each episode
        started from (``SDX_T_TARGET_INIT``), five Adam steps of BCE-with-logits on the net's ELU outputs.  Returns the last loss."""
        L = _lib.load()
        n = self.num_envs
        labels = self.env.tool_tvalue_labels(self._tv_label)
        x = self.segmentation_target_init.contiguous()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        z = None
        for _ in range(steps):
            z = self.t_value.forward(x, train=True)
            self._tv_stats.zero_()
            _lib.check(L.sdx_tvalue_bce(ctypes.c_void_p(z.data_ptr()), ctypes.c_void_p(labels.data_ptr()), n, ctypes.c_void_p(self._tv_dz.data_ptr()),
                                        ctypes.c_void_p(self._tv_stats.data_ptr()), stream))
            self.t_value.backward(self._tv_dz)
            self.t_value.adam(lr, max_norm=0.0)
        loss = self._tv_stats[0] / (2 * n)
        y = torch.nn.functional.elu(z)                                   # predict_success_confident of the last forward (TO:1321)
        self.extras["BCE_loss"] = loss
        self.extras["predict_success_confident"], self.extras["predict_unsuccess_confident"] = y[:, 0].mean(), y[:, 1].mean()
        self.extras["success_buf"] = self.success_buf[:, 0].mean()
        return loss

    def step(self, actions):
        if self.if_t_value and self.total_steps > 0 and bool(self.reset_buf.any()):     # reset_idx is about to run (TO:1446-1447)
            self.online_t_value_update()
            keep = self.success_buf.clone()           # the labels are what success_buf holds when reset_idx returns (TO:1312-1316 after :1282)
            super().step(actions)
            self.success_buf.copy_(keep)
        else:
            super().step(actions)
        self.total_steps += 1
