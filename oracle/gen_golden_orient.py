#!/usr/bin/env python
"""Golden vectors for BlockAssemblyOrient, produced by EXECUTING THE REFERENCE'S OWN PYTHON
(tasks/block_assembly/allegro_hand_block_assembly_orient.py = OR) with Isaac Gym stubbed exactly as in gen_golden.py:
    compute_observations       OR:1087-1242  (-> compute_real_observations OR:1308-1326,
                                                 compute_contact_asymmetric_observations OR:1244-1306)
    compute_reward             OR:1057-1066  (-> compute_hand_reward OR:1843-1907, TorchScript)
    pre_physics_step           OR:1697-1778  (no-reset branch)
    orientation_error / control_ik  OR:1922-1934 (through pre_physics_step)
Runs only in the build container; writes tests/golden/orient_post_physics.npz and orient_pre_physics.npz.
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
    import tasks.block_assembly.allegro_hand_block_assembly_orient as OR
    from policy_sequencing.terminal_value_function import GraspInsertTValue
    from isaacgym.torch_utils import to_torch

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from seqdex_b200.scene import Scene, quat_from_euler_zyx
    scene = Scene(task="BlockAssemblyOrient", episode_length=75, act_moving_average=0.2)
    torch.manual_seed(4321)
    rng = np.random.default_rng(4321)
    N = 24

    def rq(*shape):
        q = torch.randn(*shape, 4)
        return q / q.norm(dim=-1, keepdim=True)

    tv = GraspInsertTValue(input_dim=4, output_dim=2)
    with torch.no_grad():                           # a gate that actually switches on the random inputs below: spread the
        tv.output_layer.weight[1] *= 60.0           # logit and centre it on the > 0.99 threshold (sigmoid(4.6) = 0.99)
        tv.output_layer.bias[1] += 4.6 - float(tv(rq(4096))[:, 1].median())
    wts = torch.cat([p.detach().reshape(-1) for p in (tv.linear1.weight, tv.linear1.bias, tv.linear2.weight, tv.linear2.bias,
                                                      tv.linear3.weight, tv.linear3.bias, tv.output_layer.weight,
                                                      tv.output_layer.bias)]).numpy().astype(np.float32)

    f = Fake()
    f.num_envs, f.device = N, "cpu"
    f.gym, f.sim = mock.MagicMock(), None
    nb_env = 165
    rb = torch.zeros(N, nb_env, 13)
    rb[:, :, 0:3] = torch.randn(N, nb_env, 3) * 0.3 + torch.tensor([0.2, 0.1, 0.8])
    rb[:, :, 3:7] = rq(N, nb_env)
    rb[:, :, 7:13] = torch.randn(N, nb_env, 6) * 0.5
    rb[:, 0, 0:3] = torch.tensor([-0.35, 0.0, 0.6]); rb[:, 0, 3:7] = torch.tensor([0, 0, 0, 1.0])
    root = torch.zeros(N * 142, 13)
    root[:, 0:3] = torch.randn(N * 142, 3) * 0.2 + torch.tensor([0.25, 0.0, 0.7])
    root[:, 3:7] = rq(N * 142)
    root[:, 7:13] = torch.randn(N * 142, 6) * 0.3
    f.hand_indices = torch.arange(N) * 142
    root[f.hand_indices, 0:3] = torch.tensor([-0.35, 0.0, 0.6]); root[f.hand_indices, 3:7] = torch.tensor([0, 0, 0, 1.0])
    f.object_indices = f.hand_indices + 1
    f.extra_object_indices = f.hand_indices + 141
    seg = torch.tensor([Scene.target_brick_index(e) for e in range(N)])
    f.lego_segmentation_indices = f.hand_indices + 9 + seg
    tips = [11, 19, 23, 15]
    for e in range(0, N, 2):        # half of the envs with the fingers on the brick: the distance term of the reward vanishes
        tp = root[f.lego_segmentation_indices[e], 0:3]
        for b in tips:
            rb[e, b, 0:3] = tp + torch.randn(3) * 0.03
    for e in range(0, N, 3):        # a third nearly face up (z-align close to +1), one exactly upside down
        root[f.lego_segmentation_indices[e], 3:7] = torch.tensor([0.02, -0.03, 0.6, 0.8]) / torch.tensor([0.02, -0.03, 0.6, 0.8]).norm()
    root[f.lego_segmentation_indices[4], 3:7] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    f.root_state_tensor = root
    f.rigid_body_states = rb
    f.goal_states = torch.zeros(N, 13)
    f.hand_base_rigid_body_index = 7
    f.mount_rigid_body_index = 7
    f.fingertip_handles = torch.tensor(tips)
    f.contact_tensor = torch.randn(N, nb_env * 3) * 0.2
    f.sensor_handle_indices = torch.tensor([1, 2, 3, 4, 5, 6])
    f.envs = [None]
    f.camera_offset_quat = to_torch(quat_from_euler_zyx(0.0, -3.141 + 0.5, 1.571))       # OR:895-897 (same as GS:887-889)
    f.camera_offset_pos = to_torch([0.03, 0.107 - 0.098, 0.067 + 0.107])
    f.segmentation_target_init_pos = root[f.lego_segmentation_indices, 0:3] + torch.randn(N, 3) * 0.05
    f.segmentation_target_init_rot = rq(N)
    f.actions = torch.rand(N, 23) * 2 - 1
    f.perturb_direction = torch.zeros(N, 6)
    f.progress_buf = torch.tensor(rng.integers(0, 75, size=N), dtype=torch.long)
    f.progress_buf[0] = 73; f.progress_buf[1] = 74; f.progress_buf[2] = 75; f.progress_buf[3] = 176
    f.perturb_steps = torch.zeros(N, 1)
    f.z_unit_tensor = to_torch([0, 0, 1]).repeat(N, 1)
    f.x_unit_tensor = to_torch([1, 0, 0]).repeat(N, 1)
    f.t_value = tv
    f.obs_type = "partial_contact"
    f.save_hdf5 = False
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
    f.obs_buf_stack_frames = [prev_obs[:, 0:62].clone(), prev_obs[:, 62:124].clone(), torch.zeros(N, 62)]
    f.state_buf_stack_frames = [prev_states[:, 0:188].clone(), prev_states[:, 188:376].clone(), torch.zeros(N, 188)]
    f.compute_real_observations = lambda: OR.BlockAssemblyOrient.compute_real_observations(f)
    f.compute_contact_asymmetric_observations = lambda: OR.BlockAssemblyOrient.compute_contact_asymmetric_observations(f)
    inputs = dict(rb=rb.numpy().copy(), root=root.numpy().copy(), dof_state=dof_state.numpy().copy(), actions=f.actions.numpy().copy(),
                  init_pos=f.segmentation_target_init_pos.numpy().copy(), init_rot=f.segmentation_target_init_rot.numpy().copy(),
                  prev_obs=prev_obs.numpy(), prev_states=prev_states.numpy(), progress=f.progress_buf.numpy().copy(),
                  seg_index=seg.numpy(), tv_weights=wts)
    with torch.no_grad():
        OR.BlockAssemblyOrient.compute_observations(f)
        tv_raw = torch.sigmoid(tv(f.camera_view_segmentation_target_rot))[:, 1]
    # compute_reward (OR:1057-1066)
    f.rew_buf = torch.zeros(N)
    f.reset_buf = torch.zeros(N, dtype=torch.long); f.reset_buf[5] = 1
    inputs["reset_in"] = f.reset_buf.numpy().copy()
    f.reset_goal_buf = torch.zeros(N, dtype=torch.long)
    f.successes = torch.zeros(N); f.successes[5] = 2.0; f.successes[1] = 1.0
    inputs["successes"] = f.successes.numpy().copy()
    f.consecutive_successes = torch.tensor([0.7])
    f.spin_coef, f.hand_reset_step, f.max_episode_length = 1.0, 0, 75
    f.contacts = torch.zeros(N, 6)
    f.extra_target_pos, f.extra_target_rot = root[f.extra_object_indices, 0:3], root[f.extra_object_indices, 3:7]
    f.object_pos, f.object_rot, f.object_angvel = root[f.object_indices, 0:3], root[f.object_indices, 3:7], root[f.object_indices, 10:13]
    f.emergence_reward = torch.zeros(N); f.heap_movement_penalty = torch.zeros(N)
    f.dist_reward_scale, f.rot_reward_scale, f.rot_eps, f.action_penalty_scale = -1.0, 1.0, 0.1, -0.0
    f.success_tolerance, f.reach_goal_bonus, f.fall_dist, f.fall_penalty, f.rotation_id = 0.1, 250.0, 0.4, 0.0, 1
    f.max_consecutive_successes, f.av_factor, f.object_type = 0, to_torch(0.1), "egg"
    f.init_lego_z_align_reward = torch.zeros(N)
    f.meta_rew_buf = torch.zeros(N); f.extras = {}
    f.total_steps = 0; f.print_success_stat = False
    with torch.no_grad():
        OR.BlockAssemblyOrient.compute_reward(f, f.actions)
    np.savez(os.path.join(OUT, "orient_post_physics.npz"), obs=f.obs_buf.numpy(), states=f.states_buf.numpy(), rew=f.rew_buf.numpy(),
             reset=f.reset_buf.numpy(), tvalue=f.tvalue.detach().numpy(), tvalue_raw=tv_raw.numpy(), finger_dist=f.arm_hand_finger_dist.numpy(),
             z_align=f.lego_z_align_reward.numpy(), consec=f.consecutive_successes.numpy(), consec_in=np.array([0.7], np.float32), **inputs)
    print("orient gate:", f.tvalue.numpy().astype(int).tolist(), "rew range", float(f.rew_buf.min()), float(f.rew_buf.max()))

    # ---- pre_physics_step, no-reset branch (OR:1711-1778)
    p = Fake()
    p.num_envs, p.device = N, "cpu"
    p.gym, p.sim = mock.MagicMock(), None
    p.reset_buf = torch.zeros(N, dtype=torch.long); p.reset_goal_buf = torch.zeros(N, dtype=torch.long)
    p.test_robot_controller = False; p.use_teleoperation = False; p.apply_teleoper_perturbation = False
    p.actuated_dof_indices = torch.arange(7, 23)
    p.arm_hand_dof_lower_limits, p.arm_hand_dof_upper_limits = lo, hi
    p.act_moving_average = 0.2                                                          # yaml:16
    p.prev_targets = lo + (hi - lo) * torch.rand(N, 23)
    p.cur_targets = p.prev_targets.clone()
    p.progress_buf = torch.tensor(rng.integers(0, 75, size=N), dtype=torch.long)
    p.progress_buf[0:4] = torch.tensor([75, 76, 100, 0])
    p.segmentation_target_init_pos = torch.randn(N, 3) * 0.1 + torch.tensor([0.2, 0.2, 0.65])
    p.rigid_body_states = rb
    p.root_state_tensor = root
    p.lego_segmentation_indices = f.lego_segmentation_indices
    p.hand_base_rigid_body_index = 7
    p.target_euler = to_torch([0.0, 3.1415, 1.571]).repeat((N, 1))                      # OR:484
    jac = torch.randn(N, 23, 6, 23) * 0.4
    p.jacobian_tensor = jac
    p.arm_hand_dof_pos = dof_state[..., 0].clone()
    acts = torch.rand(N, 23) * 2 - 1
    pin = dict(prev_targets=p.prev_targets.numpy().copy(), progress=p.progress_buf.numpy().copy(),
               init_pos=p.segmentation_target_init_pos.numpy().copy(), hand_pose=rb[:, 7, 0:7].numpy().copy(),
               target_pos=root[f.lego_segmentation_indices, 0:3].numpy().copy(),
               jac7=jac[:, 6, :, :7].numpy().copy(), dof_pos=p.arm_hand_dof_pos.numpy().copy(), actions=acts.numpy().copy(),
               hand_target_quat=quat_from_euler_xyz(*p.target_euler[0]).numpy())
    OR.BlockAssemblyOrient.pre_physics_step(p, acts)
    np.savez(os.path.join(OUT, "orient_pre_physics.npz"), cur_targets=p.cur_targets.numpy(), bc_act_label=p.bc_act_label.numpy(), **pin)
    print("orient golden vectors written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
