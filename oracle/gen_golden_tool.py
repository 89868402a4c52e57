#!/usr/bin/env python
"""Golden vectors for ToolPositioningGrasp / ToolPositioningOrient, produced by EXECUTING THE REFERENCE'S OWN PYTHON
(TG = tasks/tool_positioning/allegro_hand_tool_positioning_grasp.py, TO = tasks/tool_positioning/allegro_hand_tool_positioning_orient.py)
with Isaac Gym stubbed exactly as in gen_golden.py:
    compute_observations   TG:1137-1272 / TO:1018-1135  (-> compute_contact_observations TG:1338-1368 / TO:1201-1236,
                                                            compute_contact_asymmetric_observations TG:1274-1336 / TO:1137-1199), TWICE in a
                                                            row so that the history frames are exercised
    compute_reward         TG:1077-1105 / TO:988-1016   (-> compute_hand_reward TG:1741-1893 / TO:1574-1626, TorchScript)
    pre_physics_step       TG:1580-1675 / TO:1438-1509  (no-reset branch)
    reset_idx              TG:1412-1578 (banking of good grasps, tool / hand to their start poses, history zeroed)
                           TO:1265-1436 (a banked grasp restored; a second call with `if_t_value` on: the online t-value update TO:1305-1350)
    compute_insertion_observations   tasks/tool_positioning/allegro_hand_tool_positioning_chain.py:1404-1440 (ToolPositioningChain's second buffer)
pytorch3d (third party, absent here) is needed by TG's reward for one function: oracle/p3d_transforms_restated.py.
Runs only in the build container; writes tests/golden/tool_{grasp,orient}_{post,pre,reset}.npz.

Actors per env here (what matters is only which root row is which): 0 hand, 1 object, 2 goal, 3 table, 4-8 bin boxes, 9 the tool,
10 the plate ("extra lego")."""
import os
import random
import sys
import types
from unittest import mock

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import OUT, Fake, install_stubs  # noqa: E402

NA = 11
N = 24
NB_ENV = 24 + 2 + 1 + 5 + 1 + 1      # rigid bodies per env


def main():
    os.makedirs(OUT, exist_ok=True)
    install_stubs()
    import isaacgym.torch_utils as TU
    import p3d_transforms_restated as P3D
    p3d = types.ModuleType("pytorch3d")
    p3d.transforms = P3D
    sys.modules["pytorch3d"], sys.modules["pytorch3d.transforms"] = p3d, P3D

    def quat_from_euler_xyz(roll, pitch, yaw):     # public IsaacGymEnvs torch_jit_utils restatement
        cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
        cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
        cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
        qw = cy * cr * cp + sy * sr * sp
        qx = cy * sr * cp - sy * cr * sp
        qy = cy * cr * sp + sy * sr * cp
        qz = sy * cr * cp - cy * sr * sp
        return torch.stack([qx, qy, qz, qw], dim=-1)
    TU.quat_from_euler_xyz = quat_from_euler_xyz
    if "quat_from_euler_xyz" not in TU.__all__:
        TU.__all__.append("quat_from_euler_xyz")
    import tasks.tool_positioning.allegro_hand_tool_positioning_grasp as TG
    import tasks.tool_positioning.allegro_hand_tool_positioning_orient as TO
    from isaacgym.torch_utils import to_torch

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from seqdex_b200.scene import Scene, quat_from_euler_zyx, TOOL_DEFAULT_ARM, TOOL_INSERT_PREP0
    scene = Scene("ToolPositioningGrasp")
    lo, hi = torch.from_numpy(scene.dof_lo), torch.from_numpy(scene.dof_hi)

    class Quat:                                     # gymapi.Quat().from_euler_zyx(roll, pitch, yaw)
        def from_euler_zyx(self, a, b, c):
            q = quat_from_euler_zyx(float(a), float(b), float(c))
            o = Quat(); o.x, o.y, o.z, o.w = q
            return o

    def rq(*shape):
        q = torch.randn(*shape, 4)
        return q / q.norm(dim=-1, keepdim=True)

    plate_q = torch.tensor(quat_from_euler_zyx(0.0, 3.1415, 0.0))

    for name, M, cls, ep_len in (("grasp", TG, "ToolPositioningGrasp", 150), ("orient", TO, "ToolPositioningOrient", 125)):
        C = getattr(M, cls)
        M.gymapi.Quat = Quat
        torch.manual_seed(4321 if name == "grasp" else 8765)
        rng = np.random.default_rng(4321 if name == "grasp" else 8765)
        f = Fake()
        f.num_envs, f.device = N, "cpu"
        f.gym, f.sim = mock.MagicMock(), None
        rb = torch.zeros(N, NB_ENV, 13)
        rb[:, :, 0:3] = torch.randn(N, NB_ENV, 3) * 0.3 + torch.tensor([0.2, 0.1, 0.8])
        rb[:, :, 3:7] = rq(N, NB_ENV)
        rb[:, :, 7:13] = torch.randn(N, NB_ENV, 6) * 0.5
        rb[:, 0, 0:3] = torch.tensor([-0.35, 0.0, 0.6]); rb[:, 0, 3:7] = torch.tensor([0, 0, 0, 1.0])
        root = torch.zeros(N * NA, 13)
        root[:, 0:3] = torch.randn(N * NA, 3) * 0.2 + torch.tensor([0.25, 0.0, 0.75])
        root[:, 3:7] = rq(N * NA)
        root[:, 7:13] = torch.randn(N * NA, 6) * 0.3
        f.hand_indices = torch.arange(N) * NA
        root[f.hand_indices, 0:3] = torch.tensor([-0.35, 0.0, 0.6]); root[f.hand_indices, 3:7] = torch.tensor([0, 0, 0, 1.0])
        f.object_indices = f.hand_indices + 1
        f.lego_segmentation_indices = f.hand_indices + 9
        f.extra_object_indices = f.hand_indices + 10
        root[f.extra_object_indices, 0:3] = torch.tensor([0.25, -0.2, 0.618])
        root[f.extra_object_indices, 3:7] = plate_q
        root[f.extra_object_indices, 7:13] = 0
        # a third of the tools close to the plate's orientation (bonus / small rot_dist), a few lifted
        for e in range(0, N, 3):
            dq = torch.tensor([0.05, -0.08, 0.1, 1.0]) * torch.tensor([1.0, 1.0, 1.0 + 0.3 * e, 1.0]) * torch.tensor([0.5, 0.5, 0.5, 1.0] if e == 0 else [1.0] * 4)
            root[f.lego_segmentation_indices[e], 3:7] = TU.quat_mul(plate_q[None], (dq / dq.norm())[None])[0]
        for e in range(0, N, 4):
            root[f.lego_segmentation_indices[e], 2] = 0.85 + 0.01 * e
        tips = [11, 19, 23, 15]                      # TG:218-221
        for e in range(0, N, 2):                     # half of the envs hold the tool
            tp = root[f.lego_segmentation_indices[e], 0:3]
            for b in tips:
                rb[e, b, 0:3] = tp + torch.randn(3) * 0.03
        f.root_state_tensor, f.rigid_body_states = root, rb
        f.goal_states = torch.zeros(N, 13)
        f.hand_base_rigid_body_index = f.mount_rigid_body_index = 7
        f.fingertip_handles = torch.tensor(tips)
        f.contact_tensor = torch.randn(N, NB_ENV * 3) * 0.2
        f.sensor_handle_indices = torch.tensor([1, 2, 3, 4, 5, 6])
        f.envs = [None]
        f.camera_offset_quat = to_torch(quat_from_euler_zyx(0.0, -3.141 + 0.5, 1.571))       # TG:1009-1011 (same as GS:887-889)
        f.camera_offset_pos = to_torch([0.03, 0.107 - 0.098, 0.067 + 0.107])
        init_pos = root[f.lego_segmentation_indices, 0:3] + torch.randn(N, 3) * 0.05
        init_pos[1] = root[f.lego_segmentation_indices[1], 0:3] + torch.tensor([0.09, 0.0, 0.0])     # moved out along x / y: TG:1813-1817
        init_pos[2] = root[f.lego_segmentation_indices[2], 0:3] + torch.tensor([0.0, -0.085, 0.0])
        f.segmentation_target_init_pos = init_pos
        f.segmentation_target_init_rot = rq(N)
        f.segmentation_target_init_rot[4] = torch.tensor([0.0, 0.0, 1.0, 0.0])                       # with the tool at (0, 0, 0, 1): the y axis flips under
        root[f.lego_segmentation_indices[4], 3:7] = torch.tensor([0.0, 0.0, 0.0, 1.0])               # the real-first reading -> successes = 1 (TG:1866-1868)
        f.perturb_direction = torch.zeros(N, 6)
        f.perturb_steps = torch.zeros(N, 1)
        f.obs_type = "partial_contact"
        f.arm_hand_dof_lower_limits, f.arm_hand_dof_upper_limits = lo, hi
        dof_state = torch.zeros(N, 23, 2)
        dof_state[..., 0] = lo + (hi - lo) * torch.rand(N, 23)
        dof_state[..., 1] = torch.randn(N, 23)
        f.arm_hand_dof_pos, f.arm_hand_dof_vel = dof_state[..., 0], dof_state[..., 1]
        f.vel_obs_scale, f.max_episode_length = 0.2, ep_len
        f.one_frame_num_obs, f.one_frame_num_states = 156, 188
        f.obs_buf, f.states_buf = torch.zeros(N, 468), torch.zeros(N, 564)
        hist_obs = [torch.randn(N, 156) * 0.3 for _ in range(3)]
        hist_states = [torch.randn(N, 188) * 0.3 for _ in range(3)]
        for h in hist_states:
            h[:, 141] = 0.0                          # slot 141 is never written, so no frame can hold anything there
        f.obs_buf_stack_frames = [h.clone() for h in hist_obs]
        f.state_buf_stack_frames = [h.clone() for h in hist_states]
        f.compute_contact_observations = lambda full, f=f, C=C: C.compute_contact_observations(f, full)
        f.compute_contact_asymmetric_observations = lambda f=f, C=C: C.compute_contact_asymmetric_observations(f)
        f.rew_buf = torch.zeros(N)
        f.reset_goal_buf = torch.zeros(N, dtype=torch.long)
        f.consecutive_successes = torch.tensor([0.7])
        f.spin_coef, f.hand_reset_step = 1.0, 0
        f.emergence_reward = torch.zeros(N); f.heap_movement_penalty = torch.zeros(N)
        f.dist_reward_scale, f.rot_reward_scale, f.rot_eps, f.action_penalty_scale = -1.0, 1.0, 0.1, -0.0
        f.success_tolerance, f.reach_goal_bonus, f.fall_dist, f.fall_penalty, f.rotation_id = 0.1, 250.0, 0.4, 0.0, 1
        f.max_consecutive_successes, f.av_factor, f.object_type = 0, to_torch(0.1), "egg"
        f.meta_rew_buf = torch.zeros(N); f.extras = {}
        f.total_steps = 0; f.print_success_stat = False
        f.x_unit_tensor = to_torch([1, 0, 0]).repeat((N, 1)); f.y_unit_tensor = to_torch([0, 1, 0]).repeat((N, 1))
        f.z_unit_tensor = to_torch([0, 0, 1]).repeat((N, 1))
        f.dof_force_tensor = torch.randn(N, 23)
        out = {"hist_obs": torch.stack(hist_obs[:2], 1).numpy(), "hist_states": torch.stack(hist_states[:2], 1).numpy(),
               "init_pos": f.segmentation_target_init_pos.numpy().copy(), "init_rot": f.segmentation_target_init_rot.numpy().copy(),
               "consec_in": np.array([0.7], np.float32)}
        progress = torch.tensor(rng.integers(0, ep_len - 30, size=N), dtype=torch.long)
        progress[0], progress[1], progress[2], progress[3] = ep_len - 2, 95, 50, ep_len - 1
        f.progress_buf = progress.clone()
        f.reset_buf = torch.zeros(N, dtype=torch.long); f.reset_buf[5] = 1
        f.successes = torch.zeros(N); f.successes[5] = 2.0; f.successes[1] = 1.0
        for call in (0, 1):                          # two steps: the second one's history is the first one's output
            if call == 1:                            # everything moves a little between the calls
                rb[:, :, 0:3] += torch.randn(N, NB_ENV, 3) * 0.01
                root[f.lego_segmentation_indices, 0:3] += torch.randn(N, 3) * 0.005
                dof_state[..., 0] = (dof_state[..., 0] + 0.01 * torch.randn(N, 23)).clamp(lo, hi)
                f.progress_buf += 1
            f.actions = torch.rand(N, 23) * 2 - 1
            out.update({f"rb{call}": rb.numpy().copy(), f"root{call}": root.numpy().copy(), f"dof_state{call}": dof_state.numpy().copy(),
                        f"actions{call}": f.actions.numpy().copy(), f"progress{call}": f.progress_buf.numpy().copy(),
                        f"reset_in{call}": f.reset_buf.numpy().copy(), f"successes_in{call}": f.successes.numpy().copy(),
                        f"consec_in{call}": f.consecutive_successes.numpy().copy()})
            with torch.no_grad():
                C.compute_observations(f)
                C.compute_reward(f, f.actions)
            out.update({f"obs{call}": f.obs_buf.numpy().copy(), f"states{call}": f.states_buf.numpy().copy(), f"rew{call}": f.rew_buf.numpy().copy(),
                        f"reset{call}": f.reset_buf.numpy().copy(), f"successes{call}": f.successes.numpy().copy(),
                        f"consec{call}": f.consecutive_successes.numpy().copy(), f"finger_dist{call}": f.arm_hand_finger_dist.numpy().copy()})
            print(f"tool {name} call {call}: rew range", float(f.rew_buf.min()), float(f.rew_buf.max()), "resets", int(f.reset_buf.sum()),
                  "successes", int(f.successes.sum()), "bonus envs", int((f.rew_buf > 1).sum()))
        if name == "grasp":
            # ToolPositioningChain.compute_insertion_observations (TC:1404-1440) on the state of the second call: the second observation
            # buffer the frozen inner policy reads (TC:1742)
            import tasks.tool_positioning.allegro_hand_tool_positioning_chain as TC
            f.insertion_one_frame_num_obs, f.insertion_max_episode_length = 156, 125
            f.insertion_obs_buf = torch.zeros(N, 468)
            ins_hist = [torch.randn(N, 156) * 0.3 for _ in range(3)]
            f.insertion_obs_buf_stack_frames = [h.clone() for h in ins_hist]
            f.insertion_actions = torch.rand(N, 23) * 2 - 1
            f.insertion_progress_buf = torch.tensor(rng.integers(0, 125, size=N), dtype=torch.long)
            with torch.no_grad():
                TC.ToolPositioningChain.compute_insertion_observations(f)
            out.update(ins_hist=torch.stack(ins_hist[:2], 1).numpy(), ins_actions=f.insertion_actions.numpy().copy(),
                       ins_progress=f.insertion_progress_buf.numpy().copy(), ins_obs=f.insertion_obs_buf.numpy().copy())
        np.savez(os.path.join(OUT, f"tool_{name}_post.npz"), **out)

        # ---- pre_physics_step, no-reset branch
        p = Fake()
        p.num_envs, p.device = N, "cpu"
        p.gym, p.sim = mock.MagicMock(), None
        p.reset_buf = torch.zeros(N, dtype=torch.long); p.reset_goal_buf = torch.zeros(N, dtype=torch.long)
        p.test_robot_controller = False; p.use_teleoperation = False; p.apply_teleoper_perturbation = False
        p.actuated_dof_indices = torch.arange(7, 23)
        p.arm_hand_dof_lower_limits, p.arm_hand_dof_upper_limits = lo, hi
        p.act_moving_average = 1.0
        p.prev_targets = lo + (hi - lo) * torch.rand(N, 23)
        p.prev_targets[0, 0] = hi[0] + 0.3            # TO clamps the held arm target (TO:1471-1473)
        p.cur_targets = p.prev_targets.clone()
        p.rigid_body_states = rb
        p.hand_base_rigid_body_index = 7
        p.target_euler = to_torch([0.0, 3.1415, 1.571]).repeat((N, 1))                      # TG:506
        jac = torch.randn(N, 23, 6, 23) * 0.4
        p.jacobian_tensor = jac
        p.arm_hand_dof_pos = dof_state[..., 0].clone()
        p.progress_buf = torch.tensor(rng.integers(0, 59, size=N), dtype=torch.long)
        p.progress_buf[0:6] = torch.tensor([59, 60, 61, 90, 91, 120])
        p.arm_hand_insertion_prepare_dof_pos_list = [to_torch(TOOL_INSERT_PREP0)]
        acts = torch.rand(N, 23) * 2 - 1
        pin = dict(prev_targets=p.prev_targets.numpy().copy(), hand_pose=rb[:, 7, 0:7].numpy().copy(), jac7=jac[:, 6, :, :7].numpy().copy(),
                   dof_pos=p.arm_hand_dof_pos.numpy().copy(), actions=acts.numpy().copy(), progress=p.progress_buf.numpy().copy(),
                   hand_target_quat=quat_from_euler_xyz(*p.target_euler[0]).numpy())
        C.pre_physics_step(p, acts)
        np.savez(os.path.join(OUT, f"tool_{name}_pre.npz"), cur_targets=p.cur_targets.numpy(), **pin)

        # ---- reset_idx
        r = Fake()
        r.num_envs, r.device = N, "cpu"
        r.gym, r.sim = mock.MagicMock(), None
        r.record_completion_time, r.save_hdf5, r.randomize, r.if_t_value = False, False, False, False
        r.total_steps = 11
        r.num_arm_hand_dofs = 23
        r.x_unit_tensor, r.y_unit_tensor = f.x_unit_tensor, f.y_unit_tensor
        root2 = root.clone()
        r.root_state_tensor = root2
        r.hand_indices, r.object_indices, r.extra_object_indices = f.hand_indices, f.object_indices, f.extra_object_indices
        r.goal_object_indices = f.hand_indices + 2
        r.lego_indices = (f.hand_indices[:, None] + 9).long()
        r.lego_segmentation_indices = f.lego_segmentation_indices.clone()
        r.pre_exchange_lego_segmentation_indices = f.lego_segmentation_indices.clone()
        r.segmentation_target_rot, r.segmentation_target_pos = root2[f.lego_segmentation_indices, 3:7].clone(), root2[f.lego_segmentation_indices, 0:3].clone()
        r.extra_target_rot, r.extra_target_pos = root2[f.extra_object_indices, 3:7].clone(), root2[f.extra_object_indices, 0:3].clone()
        r.arm_hand_finger_dist = torch.tensor(rng.uniform(0.1, 0.7, size=N), dtype=torch.float32)
        r.success_buf = torch.zeros(N, 2)
        r.rigid_body_states = rb.clone()
        r.base_pos = r.rigid_body_states[:, 0, 0:3]
        r.rb_forces = torch.zeros(N, NB_ENV, 3)
        r.object_init_state = torch.zeros(N, 13); r.object_init_state[:, 0:3] = torch.tensor([0.0, 0.0, -10.78]); r.object_init_state[:, 6] = 1
        r.goal_states = r.object_init_state.clone(); r.goal_init_state = r.object_init_state.clone()
        r.goal_displacement_tensor = torch.tensor([-0.2, -0.06, 0.12])
        r.reset_goal_buf = torch.zeros(N, dtype=torch.long)
        r.reset_position_noise, r.up_axis_idx = 0.0, 2
        r.object_pose_for_open_loop = torch.zeros(N, 7)
        lego_init = torch.zeros(N, 1, 13)
        lego_init[:, 0, 0:3] = torch.tensor([1.08, 0.08, 0.62]); lego_init[:, 0, 3:7] = torch.tensor(quat_from_euler_zyx(0.0, 0.0, 0.785))
        r.lego_init_states = lego_init
        dof2 = dof_state.clone()
        r.dof_state = dof2.view(N * 23, 2)
        r.arm_hand_dof_pos, r.arm_hand_dof_vel = dof2[..., 0], dof2[..., 1]
        r.arm_hand_dof_lower_limits, r.arm_hand_dof_upper_limits = lo, hi
        r.prev_targets, r.cur_targets = torch.randn(N, 23), torch.randn(N, 23)
        r.t_value_obs_buf = torch.zeros(N, 7)
        r.random_force_prob = torch.zeros(N); r.force_prob_range = to_torch([0.001, 0.1])
        r.segmentation_target_init_pos, r.segmentation_target_init_rot = torch.zeros(N, 3), torch.zeros(N, 4)
        r.progress_buf = torch.tensor(rng.integers(1, ep_len, size=N), dtype=torch.long)
        r.reset_buf = torch.zeros(N, dtype=torch.long)
        env_ids = torch.tensor([0, 3, 4, 5, 8, 9, 12, 13, 14, 16, 20, 22])
        r.reset_buf[env_ids] = 1
        r.successes = torch.rand(N); r.meta_rew_buf = torch.rand(N)
        r.perturb_steps = torch.zeros(N); r.perturb_direction = torch.zeros(N, 6)
        r.reset_target_pose = lambda ids, apply_reset=False, r=r, C=C: C.reset_target_pose(r, ids, apply_reset)
        rin = dict(root=root2.numpy().copy(), dof_state=dof2.numpy().copy(), env_ids=env_ids.numpy(), progress=r.progress_buf.numpy().copy(),
                   successes=r.successes.numpy().copy(), finger_dist=r.arm_hand_finger_dist.numpy().copy())
        extra = {}
        if name == "grasp":
            arm = torch.zeros(23); arm[:7] = torch.tensor(TOOL_DEFAULT_ARM)
            arm[7:] = TU.scale(torch.ones(16), lo[7:], hi[7:])                                   # TG:283-290
            r.arm_hand_default_dof_pos = arm
            r.arm_hand_dof_default_vel = torch.zeros(23)
            r.contact_obs_buf = torch.randn(N, 30)
            r.obs_buf = torch.randn(N, 468)
            r.obs_buf_stack_frames = [torch.randn(N, 156) for _ in range(3)]
            r.state_buf_stack_frames = [torch.randn(N, 188) for _ in range(3)]
            # the reference's eight lists alias ONE tensor (TG:441-442); separate tensors here, which is what the engine's rings are
            r.saved_grasp_hand_ternimal_states_list = [torch.zeros(10000 + 1024, 23, 2) for _ in range(8)]
            r.saved_grasp_object_ternimal_states_list = [torch.zeros(10000 + 1024, 13) for _ in range(8)]
            r.saved_grasp_ternimal_states_index_list = [0, 3, 9999, 10000, 7, 0, 1, 2]
            r.saved_orient_grasp_init_index_list = [0] * 8
            rin["index_in"] = np.asarray(r.saved_grasp_ternimal_states_index_list)
            # make the banking gate pass for some of the resetting envs: lifted, fingers close, near the plate's orientation (rot_dist < 1)
            for e in (0, 3, 8, 9, 12, 16, 20):
                root2[f.lego_segmentation_indices[e], 2] = 0.82 + 0.002 * e
                r.arm_hand_finger_dist[e] = 0.2 + 0.005 * e
                dq = torch.tensor([0.1, -0.2, 0.05 * (e % 5), 1.0]) * torch.tensor([0.4, 0.4, 0.4, 1.0] if e in (0, 9) else [1.0] * 4)
                root2[f.lego_segmentation_indices[e], 3:7] = TU.quat_mul(plate_q[None], (dq / dq.norm())[None])[0]
            root2[f.lego_segmentation_indices[16], 2] = 0.79        # too low
            r.arm_hand_finger_dist[20] = 0.45                         # fingers too far
            r.segmentation_target_rot, r.segmentation_target_pos = root2[f.lego_segmentation_indices, 3:7].clone(), root2[f.lego_segmentation_indices, 0:3].clone()
            rin["root"], rin["finger_dist"] = root2.numpy().copy(), r.arm_hand_finger_dist.numpy().copy()
            draws = {}
            real_rand = M.torch_rand_float

            def rec_rand(lower, upper, shape, device):
                x = real_rand(lower, upper, shape, device)
                if shape[1] == 23 * 2 + 5:
                    draws["u"] = x[:, 5].clone()
                return x
            pitch_k = 3
            with mock.patch.object(M, "torch_rand_float", rec_rand), mock.patch.object(M.random, "sample", lambda pop, k: [pitch_k]), \
                    mock.patch.object(M, "print", lambda *a, **k: None, create=True):
                C.reset_idx(r, env_ids, torch.tensor([], dtype=torch.long))
            u = torch.zeros(N); u[env_ids] = draws["u"]
            bh, bo = torch.stack(r.saved_grasp_hand_ternimal_states_list).numpy(), torch.stack(r.saved_grasp_object_ternimal_states_list).numpy()
            where = np.argwhere(np.abs(bo).sum(-1) > 0)             # the (type, slot) pairs that were written; every other row is still zero
            extra = dict(yaw_u=u.numpy(), pitch_k=np.int64(pitch_k), obs_out=r.obs_buf.numpy(),
                         frames_obs=torch.stack(r.obs_buf_stack_frames, 1).numpy(), frames_states=torch.stack(r.state_buf_stack_frames, 1).numpy(),
                         bank_where=where, bank_hand_rows=bh[where[:, 0], where[:, 1]], bank_obj_rows=bo[where[:, 0], where[:, 1]],
                         index_out=np.asarray(r.saved_grasp_ternimal_states_index_list))
            print("tool grasp reset: ring indices", rin["index_in"].tolist(), "->", extra["index_out"].tolist())
        else:
            PER = 6
            r.saved_grasping_object_ternimal_states_list = [torch.cat([torch.randn(PER, 1, 3) * 0.05 + torch.tensor([0.2, -0.1, 0.9]), rq(PER, 1), torch.randn(PER, 1, 6)], -1)
                                                            for _ in range(8)]
            r.saved_grasping_hand_ternimal_states_list = [torch.stack([lo + (hi - lo) * torch.rand(PER, 23), torch.randn(PER, 23)], -1) for _ in range(8)]
            slots = [int(x) for x in rng.integers(0, PER, size=len(env_ids))]
            calls = {"n": 0}
            real_sample = random.sample

            def fake_sample(pop, k):
                if isinstance(pop, range) and len(pop) == 5000:
                    s = slots[calls["n"]]; calls["n"] += 1
                    return [s]
                return real_sample(pop, k)
            random.seed(5)
            with mock.patch.object(M.random, "sample", fake_sample), mock.patch.object(M, "print", lambda *a, **k: None, create=True):
                C.reset_idx(r, env_ids, torch.tensor([], dtype=torch.long))
            so = np.zeros(N, np.int32); so[env_ids.numpy()] = slots
            extra = dict(slot_by_env=so, bank_obj=torch.stack(r.saved_grasping_object_ternimal_states_list).numpy(),
                         bank_hand=torch.stack(r.saved_grasping_hand_ternimal_states_list).numpy(), t_value_obs=r.t_value_obs_buf.numpy().copy())
            first = dict(root_out=root2.numpy().copy(), dof_out=dof2.numpy().copy(), prev_targets=r.prev_targets.numpy().copy(),
                         cur_targets=r.cur_targets.numpy().copy(), init_pos=r.segmentation_target_init_pos.numpy().copy(),
                         init_rot=r.segmentation_target_init_rot.numpy().copy(), progress_out=r.progress_buf.numpy().copy(),
                         reset_out=r.reset_buf.numpy().copy(), successes_out=r.successes.numpy().copy(), success_buf=r.success_buf.numpy().copy())
            # ---- a second reset_idx call with the online t-value update switched on (TO:1305-1350; `if_t_value` is hard-wired False at TO:377)
            from policy_sequencing.terminal_value_function import GraspInsertTValue
            torch.manual_seed(99)
            r.if_t_value = True
            r.t_value = GraspInsertTValue(input_dim=7, output_dim=2)
            r.t_value_optimizer = torch.optim.Adam(r.t_value.parameters(), lr=0.0003)              # TO:384
            r.bce_logits_loss = torch.nn.BCEWithLogitsLoss()                                       # TO:387
            r.extras, r.max_episode_length, r.t_value_save_path = {}, 125, "/nonexistent"
            r.t_value_obs_buf = torch.cat([torch.randn(N, 3) * 0.1 + torch.tensor([0.2, -0.1, 0.9]), rq(N)], -1)
            r.extra_target_pos, r.extra_target_rot = root2[f.extra_object_indices, 0:3].clone(), root2[f.extra_object_indices, 3:7].clone()
            r.symmetry_extra_target_rot = TU.quat_mul(r.extra_target_rot, to_torch([0.0, 0.0, 1.0, 0.0]).repeat(N, 1))
            tpos, trot = root2[f.lego_segmentation_indices, 0:3].clone(), root2[f.lego_segmentation_indices, 3:7].clone()
            for e in range(0, N, 2):                    # half of the envs ended aligned with the plate (or its pi-about-z twin) within 1 cm
                tpos[e] = r.extra_target_pos[e] + torch.randn(3) * 0.003
                dq = torch.tensor([0.01, -0.02, 0.015, 1.0])
                base = r.symmetry_extra_target_rot[e] if e % 4 == 0 else r.extra_target_rot[e]
                trot[e] = TU.quat_mul(base[None], (dq / dq.norm())[None])[0]
            tpos[2] = r.extra_target_pos[2] + torch.tensor([0.02, 0.0, 0.0])                       # aligned but 2 cm away: a failure
            r.segmentation_target_pos, r.segmentation_target_rot = tpos, trot
            w0 = torch.cat([p_.detach().reshape(-1) for p_ in r.t_value.parameters()]).numpy().copy()
            losses = []
            real_bce = r.bce_logits_loss

            def rec_bce(pred, target):
                out_ = real_bce(pred, target)
                losses.append(float(out_))
                return out_
            r.bce_logits_loss = rec_bce
            env_ids2 = torch.tensor([1, 2, 6])
            r.reset_buf[env_ids2] = 1
            tv_in = r.t_value_obs_buf.numpy().copy()
            calls["n"] = 0
            slots[:] = slots[:3] + slots[3:]
            with mock.patch.object(M.random, "sample", fake_sample), mock.patch.object(M, "print", lambda *a, **k: None, create=True):
                C.reset_idx(r, env_ids2, torch.tensor([], dtype=torch.long))
            w5 = torch.cat([p_.detach().reshape(-1) for p_ in r.t_value.parameters()]).numpy().copy()
            extra.update(tv_obs_in=tv_in, tv_target_pos=tpos.numpy(), tv_target_rot=trot.numpy(), tv_plate=root2[f.extra_object_indices, 0:7].numpy().copy(),
                         tv_success_buf=r.success_buf.numpy().copy(), tv_w0=w0, tv_w5=w5, tv_losses=np.asarray(losses, np.float32),
                         tv_pred_last=r.predict_success_confident.detach().numpy().copy())
            print("tool orient online t-value: successes", int(r.success_buf[:, 0].sum()), "of", N, "losses", [round(x, 4) for x in losses])
        outs = first if name == "orient" else dict(
            root_out=root2.numpy(), dof_out=dof2.numpy(), prev_targets=r.prev_targets.numpy(), cur_targets=r.cur_targets.numpy(),
            init_pos=r.segmentation_target_init_pos.numpy(), init_rot=r.segmentation_target_init_rot.numpy(), progress_out=r.progress_buf.numpy(),
            reset_out=r.reset_buf.numpy(), successes_out=r.successes.numpy(), success_buf=r.success_buf.numpy())
        np.savez(os.path.join(OUT, f"tool_{name}_reset.npz"), **outs, **rin, **extra)
        print(f"tool {name} golden vectors written to", os.path.normpath(OUT), "| successes at reset:", outs["success_buf"][env_ids.numpy(), 0].tolist())


if __name__ == "__main__":
    main()
