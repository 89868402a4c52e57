#!/usr/bin/env python
"""Golden vectors for BlockAssemblySearch, produced by EXECUTING THE REFERENCE'S OWN PYTHON
(tasks/block_assembly/allegro_hand_block_assembly_search.py = SE) with Isaac Gym stubbed exactly as in gen_golden.py:
    compute_observations       SE:984-1166   (camera branch off; -> compute_contact_observations SE:1220-1245 on synthetic
                                                segmentation images, compute_contact_asymmetric_observations SE:1168-1218)
    compute_reward             SE:944-952    (-> compute_hand_reward SE:1660-1712)
    pre_physics_step           SE:1539-1596  (no-reset branch)
Runs only in the build container; writes tests/golden/search_post_physics.npz and search_pre_physics.npz.
"""
import os
import sys
from unittest import mock

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import OUT, Fake, install_stubs  # noqa: E402


def main():
    os.makedirs(OUT, exist_ok=True)
    install_stubs()
    import isaacgym.torch_utils as TU

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
    TU.__all__.append("quat_from_euler_xyz")
    import tasks.block_assembly.allegro_hand_block_assembly_search as SE
    from policy_sequencing.terminal_value_function import RetriGraspTValue
    from isaacgym.torch_utils import to_torch

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from seqdex_b200.scene import Scene, quat_from_euler_zyx
    scene = Scene(task="BlockAssemblySearch", episode_length=75, act_moving_average=0.6)
    torch.manual_seed(777)
    rng = np.random.default_rng(777)
    N = 24

    def rq(*shape):
        q = torch.randn(*shape, 4)
        return q / q.norm(dim=-1, keepdim=True)

    f = Fake()
    f.num_envs, f.device = N, "cpu"
    f.gym, f.sim = mock.MagicMock(), None
    f.enable_camera_sensors = False
    nb_env = 165
    rb = torch.zeros(N, nb_env, 13)
    rb[:, :, 0:3] = torch.randn(N, nb_env, 3) * 0.3 + torch.tensor([0.2, 0.1, 0.8])
    rb[:, :, 3:7] = rq(N, nb_env)
    rb[:, :, 7:13] = torch.randn(N, nb_env, 6) * 0.5
    root = torch.zeros(N * 142, 13)
    root[:, 0:3] = torch.randn(N * 142, 3) * 0.2 + torch.tensor([0.25, 0.0, 0.7])
    root[:, 3:7] = rq(N * 142)
    root[:, 7:13] = torch.randn(N * 142, 6) * 0.3
    f.hand_indices = torch.arange(N) * 142
    root[f.hand_indices, 0:3] = torch.tensor([-0.35, 0.0, 0.6]); root[f.hand_indices, 3:7] = torch.tensor([0, 0, 0, 1.0])
    f.object_indices = f.hand_indices + 1
    seg = torch.tensor([Scene.target_brick_index(e) for e in range(N)])
    f.lego_segmentation_indices = f.hand_indices + 9 + seg
    tips = [11, 19, 23, 15]
    for e in range(0, N, 2):
        tp = root[f.lego_segmentation_indices[e], 0:3]
        for b in tips:
            rb[e, b, 0:3] = tp + torch.randn(3) * 0.03
    f.root_state_tensor = root
    f.rigid_body_states = rb
    f.goal_states = torch.zeros(N, 13)
    f.hand_base_rigid_body_index = 7
    f.mount_rigid_body_index = 7
    f.fingertip_handles = torch.tensor(tips)
    contact = torch.randn(N, nb_env * 3) * 0.08                      # norms straddle the 0.1 N threshold (SE:1115)
    f.contact_tensor = contact
    f.sensor_handle_indices = to_torch([0, 1, 2, 3, 4, 5, 6], dtype=torch.int64)      # SE:919-920
    f.envs = [None]
    f.camera_offset_quat = to_torch(quat_from_euler_zyx(0.0, -3.141 + 0.5, 1.571))   # SE:795-797 (same as GS:887-889)
    f.camera_offset_pos = to_torch([0.03, 0.107 - 0.098, 0.067 + 0.107])
    f.segmentation_target_init_pos = root[f.lego_segmentation_indices, 0:3] + torch.randn(N, 3) * 0.05
    f.segmentation_target_init_rot = rq(N)
    f.actions = torch.rand(N, 23) * 2 - 1
    f.perturb_direction = torch.zeros(N, 6)
    f.progress_buf = torch.tensor(rng.integers(1, 73, size=N), dtype=torch.long)
    f.progress_buf[1] = 74; f.progress_buf[2] = 75; f.progress_buf[3] = 73
    f.max_episode_length = 75
    f.hand_reset_step = 45
    f.perturb_steps = torch.zeros(N, 1)
    f.hand_pos_history = torch.zeros(N, 45 * 8 + 1, 3)
    for k in range(8):
        setattr(f, f"hand_pos_history_{k}", torch.zeros(N, 3))                        # SE:1457-1465: means of a zeroed buffer
    tv = RetriGraspTValue(input_dim=650, output_dim=2)
    f.t_value = tv
    prev_tvobs = torch.randn(N, 650) * 0.2
    f.t_value_obs_buf = prev_tvobs.clone()
    f.obs_type, f.asymmetric_obs, f.save_hdf5 = "partial_contact", True, False
    lo, hi = torch.from_numpy(scene.dof_lo), torch.from_numpy(scene.dof_hi)
    f.arm_hand_dof_lower_limits, f.arm_hand_dof_upper_limits = lo, hi
    dof_state = torch.zeros(N, 23, 2)
    dof_state[..., 0] = lo + (hi - lo) * torch.rand(N, 23)
    dof_state[..., 1] = torch.randn(N, 23)
    f.arm_hand_dof_pos, f.arm_hand_dof_vel = dof_state[..., 0], dof_state[..., 1]
    f.vel_obs_scale = 0.2
    f.one_frame_num_obs, f.one_frame_num_states = 62, 188
    prev_obs = torch.randn(N, 186) * 0.3
    prev_states = torch.randn(N, 564) * 0.3
    f.obs_buf, f.states_buf = prev_obs.clone(), prev_states.clone()
    # synthetic segmentation images: a blob of the target's id (lego_i + 1, SE:846-847) among other ids; some envs see nothing
    f.segmentation_id_list = [int(s) + 1 for s in seg]
    imgs = []
    for e in range(N):
        img = torch.tensor(rng.integers(0, 9, size=(128, 128)), dtype=torch.int32)
        img[img == f.segmentation_id_list[e]] = 0
        if e % 5 != 4:
            r0, c0, h, w = (int(v) for v in (rng.integers(0, 100), rng.integers(0, 100), rng.integers(1, 28), rng.integers(1, 28)))
            blob = torch.tensor(rng.uniform(size=(h, w)) < 0.8)
            img[r0:r0 + h, c0:c0 + w][blob] = f.segmentation_id_list[e]
        imgs.append(img)
    f.camera_seg_tensors = imgs
    f.segmentation_object_center_point_x = torch.zeros((N, 1), dtype=torch.int)
    f.segmentation_object_center_point_y = torch.zeros((N, 1), dtype=torch.int)
    f.segmentation_object_point_num = torch.zeros((N, 1), dtype=torch.int)
    f.compute_contact_observations = lambda full: SE.BlockAssemblySearch.compute_contact_observations(f, full)
    f.compute_contact_asymmetric_observations = lambda: SE.BlockAssemblySearch.compute_contact_asymmetric_observations(f)
    inputs = dict(rb=rb.numpy().copy(), root=root.numpy().copy(), dof_state=dof_state.numpy().copy(), actions=f.actions.numpy().copy(),
                  init_pos=f.segmentation_target_init_pos.numpy().copy(), init_rot=f.segmentation_target_init_rot.numpy().copy(),
                  prev_obs=prev_obs.numpy(), prev_states=prev_states.numpy(), prev_tvobs=prev_tvobs.numpy(),
                  progress=f.progress_buf.numpy().copy(), seg_index=seg.numpy(), contact=contact.numpy().reshape(N, nb_env, 3)[:, :24].copy(),
                  masks=np.stack([(imgs[e] == f.segmentation_id_list[e]).numpy() for e in range(N)]).astype(np.uint8))
    with torch.no_grad():
        SE.BlockAssemblySearch.compute_observations(f)
    segf = torch.cat([f.segmentation_object_point_num, f.segmentation_object_center_point_x, f.segmentation_object_center_point_y], dim=1)
    # compute_reward (SE:944-952)
    f.rew_buf = torch.zeros(N)
    f.reset_buf = torch.zeros(N, dtype=torch.long); f.reset_buf[5] = 1
    inputs["reset_in"] = f.reset_buf.numpy().copy()
    f.reset_goal_buf = torch.zeros(N, dtype=torch.long)
    f.successes = torch.zeros(N); f.successes[5] = 2.0; f.successes[1] = 1.0
    inputs["successes"] = f.successes.numpy().copy()
    f.consecutive_successes = torch.tensor([0.7])
    f.spin_coef = 1.0
    f.goal_pos, f.goal_rot = f.goal_states[:, 0:3], f.goal_states[:, 3:7]
    f.emergence_reward = torch.randn(N)
    f.heap_movement_penalty = torch.zeros(N); f.init_heap_movement_penalty = torch.zeros(N)
    f.dist_reward_scale, f.rot_reward_scale, f.rot_eps, f.action_penalty_scale = -1.0, 1.0, 0.1, -0.0
    f.success_tolerance, f.reach_goal_bonus, f.fall_dist, f.fall_penalty, f.rotation_id = 0.1, 250.0, 0.4, 0.0, 1
    f.max_consecutive_successes, f.av_factor, f.object_type = 0, to_torch(0.1), "egg"
    f.meta_rew_buf = torch.zeros(N); f.extras = {}
    f.total_steps = 0; f.print_success_stat = False
    with torch.no_grad():
        SE.BlockAssemblySearch.compute_reward(f, f.actions)
    np.savez(os.path.join(OUT, "search_post_physics.npz"), obs=f.obs_buf.numpy(), states=f.states_buf.numpy(), rew=f.rew_buf.numpy(),
             reset=f.reset_buf.numpy(), seg=segf.numpy().astype(np.int32), tvobs=f.t_value_obs_buf.numpy(),
             finger_dist_states=f.states_buf.numpy()[:, 0], consec=f.consecutive_successes.numpy(), consec_in=np.array([0.7], np.float32),
             contacts=f.contacts.numpy(), **inputs)
    print("search seg features:", segf[:6].tolist(), "rew range", float(f.rew_buf.min()), float(f.rew_buf.max()), "contacts", f.contacts.sum(-1)[:8].tolist())

    # ---- pre_physics_step, no-reset branch (SE:1546-1596)
    p = Fake()
    p.num_envs, p.device = N, "cpu"
    p.gym, p.sim = mock.MagicMock(), None
    p.reset_buf = torch.zeros(N, dtype=torch.long); p.reset_goal_buf = torch.zeros(N, dtype=torch.long)
    p.test_robot_controller = False; p.apply_teleoper_perturbation = False
    p.actuated_dof_indices = torch.arange(7, 23)
    p.arm_hand_dof_lower_limits, p.arm_hand_dof_upper_limits = lo, hi
    p.act_moving_average = 0.6                                                          # yaml:16
    p.prev_targets = lo + (hi - lo) * torch.rand(N, 23)
    p.cur_targets = p.prev_targets.clone()
    p.rigid_body_states = rb
    p.segmentation_target_pos = root[f.lego_segmentation_indices, 0:3].clone()
    p.hand_base_rigid_body_index = 7
    jac = torch.randn(N, 23, 6, 23) * 0.4
    p.jacobian_tensor = jac
    p.arm_hand_dof_pos = dof_state[..., 0].clone()
    acts = (torch.rand(N, 23) * 2 - 1) * 1.0
    pin = dict(prev_targets=p.prev_targets.numpy().copy(), hand_pose=rb[:, 7, 0:7].numpy().copy(),
               target_pos=p.segmentation_target_pos.numpy().copy(), jac7=jac[:, 6, :, :7].numpy().copy(),
               dof_pos=p.arm_hand_dof_pos.numpy().copy(), actions=acts.numpy().copy(),
               hand_target_quat=quat_from_euler_xyz(*to_torch([0.0, 3.14, 1.57])).numpy())
    SE.BlockAssemblySearch.pre_physics_step(p, acts)
    np.savez(os.path.join(OUT, "search_pre_physics.npz"), cur_targets=p.cur_targets.numpy(), **pin)
    print("search golden vectors written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
