#!/usr/bin/env python
"""Golden vectors for BlockAssemblyInsertSim, produced by EXECUTING THE REFERENCE'S OWN PYTHON
(tasks/block_assembly/allegro_hand_block_assembly_insert_sim.py = IS) with Isaac Gym stubbed exactly as in gen_golden.py:
    compute_observations       IS:1090-1220  (-> compute_contact_observations IS:1280-1298,
                                                 compute_contact_asymmetric_observations IS:1222-1278)
    compute_reward             IS:1057-1066  (-> compute_hand_reward IS:1640-1694, TorchScript)
    pre_physics_step           IS:1495-1565  (no-reset branch; orientation_error / control_ik IS:1712-1725)
    reset_idx                  IS:1328-1493  (restores a banked grasp: target-brick root row + hand DoF state, IS:1449-1453)
Runs only in the build container; writes tests/golden/insert_post_physics.npz, insert_pre_physics.npz, insert_reset.npz.

Actors per env here (what matters is only which root row is which): 0 hand, 1 object, 2 goal, 3 table, 4-8 bin boxes, 9-16 the
eight bricks (one per type, IS:689-736), 17 the base-plate ("extra lego", 4x4x{1,2,4} by env % 3, IS:971-977).
"""
import os
import random
import sys
from unittest import mock

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import OUT, Fake, install_stubs  # noqa: E402

NA = 18            # actors per env in this stand-in
N = 24


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
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath("/root/reference/dexteroushandenvs/x"))))   # IS:389 imports dexteroushandenvs.policy_sequencing...
    import tasks.block_assembly.allegro_hand_block_assembly_insert_sim as IS
    from isaacgym.torch_utils import to_torch

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from seqdex_b200.scene import Scene, quat_from_euler_zyx
    scene = Scene()
    torch.manual_seed(9876)
    rng = np.random.default_rng(9876)

    def rq(*shape):
        q = torch.randn(*shape, 4)
        return q / q.norm(dim=-1, keepdim=True)

    def index_lists(f):
        """IS:786-812: which envs get which target offsets (use_unseen False)"""
        f.extra_1xn_lego_pos_offset_indices = [i for i in range(N) if i % 8 in (0, 1, 2, 6, 3, 4, 7)]
        f.extra_1x1_lego_pos_offset_indices = [i for i in range(N) if i % 8 == 5]
        f.extra_height_lego_pos_offset_indices_0 = [i for i in range(N) if i % 3 == 0]
        f.extra_height_lego_pos_offset_indices_1 = [i for i in range(N) if i % 3 == 1]
        f.extra_height_lego_pos_offset_indices_2 = [i for i in range(N) if i % 3 == 2]

    f = Fake()
    f.num_envs, f.device = N, "cpu"
    f.gym, f.sim = mock.MagicMock(), None
    nb_env = 24 + 2 + 1 + 5 + 8 + 1
    rb = torch.zeros(N, nb_env, 13)
    rb[:, :, 0:3] = torch.randn(N, nb_env, 3) * 0.3 + torch.tensor([0.2, 0.1, 0.8])
    rb[:, :, 3:7] = rq(N, nb_env)
    rb[:, :, 7:13] = torch.randn(N, nb_env, 6) * 0.5
    rb[:, 0, 0:3] = torch.tensor([-0.35, 0.0, 0.6]); rb[:, 0, 3:7] = torch.tensor([0, 0, 0, 1.0])
    root = torch.zeros(N * NA, 13)
    root[:, 0:3] = torch.randn(N * NA, 3) * 0.2 + torch.tensor([0.25, 0.0, 0.7])
    root[:, 3:7] = rq(N * NA)
    root[:, 7:13] = torch.randn(N * NA, 6) * 0.3
    f.hand_indices = torch.arange(N) * NA
    root[f.hand_indices, 0:3] = torch.tensor([-0.35, 0.0, 0.6]); root[f.hand_indices, 3:7] = torch.tensor([0, 0, 0, 1.0])
    f.object_indices = f.hand_indices + 1
    f.extra_object_indices = f.hand_indices + 17
    seg = torch.tensor([Scene.target_brick_index(e) for e in range(N)])         # IS:941-943: env % 8, {3, 4, 7} -> 0
    f.lego_segmentation_indices = f.hand_indices + 9 + seg
    # base-plates near their place, upright up to a yaw; a third of the bricks close to the insertion pose
    yaw = torch.tensor(rng.uniform(-3.1, 3.1, size=N), dtype=torch.float32)
    root[f.extra_object_indices, 0:3] = torch.tensor([0.25, -0.2, 0.618]) + torch.randn(N, 3) * 0.01
    root[f.extra_object_indices, 3:7] = torch.stack([torch.zeros(N), torch.zeros(N), torch.sin(yaw / 2), torch.cos(yaw / 2)], -1)
    index_lists(f)
    tips = [11, 19, 23, 15]
    for e in range(0, N, 2):        # half of the envs hold the brick (finger distance below the 0.6 reset threshold)
        tp = root[f.lego_segmentation_indices[e], 0:3]
        for b in tips:
            rb[e, b, 0:3] = tp + torch.randn(3) * 0.03
    f.root_state_tensor = root
    f.rigid_body_states = rb
    f.goal_states = torch.zeros(N, 13)
    f.hand_base_rigid_body_index = 7
    f.mount_rigid_body_index = 7
    f.fingertip_handles = torch.tensor(tips)
    f.contact_tensor = torch.randn(N, nb_env * 3) * 0.2
    f.sensor_handle_indices = torch.tensor([1, 2, 3, 4, 5, 6])
    f.envs = [None]
    f.camera_offset_quat = to_torch(quat_from_euler_zyx(0.0, -3.141 + 0.5, 1.571))       # IS:880-882 (same as GS:887-889)
    f.camera_offset_pos = to_torch([0.03, 0.107 - 0.098, 0.067 + 0.107])
    f.segmentation_target_init_pos = root[f.lego_segmentation_indices, 0:3] + torch.randn(N, 3) * 0.05
    f.segmentation_target_init_rot = rq(N)
    f.actions = torch.rand(N, 23) * 2 - 1
    f.perturb_direction = torch.zeros(N, 6)
    f.progress_buf = torch.tensor(rng.integers(0, 120, size=N), dtype=torch.long)
    f.progress_buf[0] = 123; f.progress_buf[1] = 124; f.progress_buf[2] = 125
    f.perturb_steps = torch.zeros(N, 1)
    f.obs_type = "partial_contact"
    f.save_hdf5 = False
    lo, hi = torch.from_numpy(scene.dof_lo), torch.from_numpy(scene.dof_hi)
    f.arm_hand_dof_lower_limits, f.arm_hand_dof_upper_limits = lo, hi
    dof_state = torch.zeros(N, 23, 2)
    dof_state[..., 0] = lo + (hi - lo) * torch.rand(N, 23)
    dof_state[..., 1] = torch.randn(N, 23)
    f.arm_hand_dof_pos, f.arm_hand_dof_vel = dof_state[..., 0], dof_state[..., 1]
    f.vel_obs_scale, f.max_episode_length = 0.2, 125
    prev_obs, prev_states = torch.randn(N, 75) * 0.3, torch.randn(N, 188) * 0.3
    f.obs_buf, f.states_buf = prev_obs.clone(), prev_states.clone()
    f.compute_contact_observations = lambda full: IS.BlockAssemblyInsertSim.compute_contact_observations(f, full)
    f.compute_contact_asymmetric_observations = lambda: IS.BlockAssemblyInsertSim.compute_contact_asymmetric_observations(f)
    # put the brick of every third env AT the insertion pose (success bonus, small rotation error); the pose the reference computes:
    # plate position + R (0, 0, 0.0375 h) + R (0, 0.015, 0) [+ R (0.015, 0, 0) for the 1x1] (IS:1124-1132)
    for e in range(0, N, 3):
        q = root[f.extra_object_indices[e], 3:7]
        h = [1, 2, 3][e % 3]
        off = torch.tensor([0.015 if e % 8 == 5 else 0.0, 0.015, 0.0375 * h])
        root[f.lego_segmentation_indices[e], 0:3] = root[f.extra_object_indices[e], 0:3] + TU.quat_apply(q[None], off[None])[0] + torch.randn(3) * 0.004
        root[f.lego_segmentation_indices[e], 3:7] = TU.quat_mul(q[None], (torch.tensor([0.02, -0.01, 0.03, 1.0]) / torch.tensor([0.02, -0.01, 0.03, 1.0]).norm())[None])[0]
    root[f.lego_segmentation_indices[3], 3:7] = TU.quat_mul(root[f.extra_object_indices[3], 3:7][None], torch.tensor([[0.0, 0.0, 1.0, 0.0]]))[0]   # the symmetric pose
    inputs = dict(rb=rb.numpy().copy(), root=root.numpy().copy(), dof_state=dof_state.numpy().copy(), actions=f.actions.numpy().copy(),
                  init_pos=f.segmentation_target_init_pos.numpy().copy(), init_rot=f.segmentation_target_init_rot.numpy().copy(),
                  prev_obs=prev_obs.numpy(), prev_states=prev_states.numpy(), progress=f.progress_buf.numpy().copy(), seg_index=seg.numpy())
    with torch.no_grad():
        IS.BlockAssemblyInsertSim.compute_observations(f)
    # compute_reward (IS:1057-1066)
    f.rew_buf = torch.zeros(N)
    f.reset_buf = torch.zeros(N, dtype=torch.long); f.reset_buf[5] = 1
    inputs["reset_in"] = f.reset_buf.numpy().copy()
    f.reset_goal_buf = torch.zeros(N, dtype=torch.long)
    f.successes = torch.zeros(N); f.successes[5] = 2.0; f.successes[1] = 1.0
    inputs["successes"] = f.successes.numpy().copy()
    f.consecutive_successes = torch.tensor([0.7])
    f.rot_err = torch.randn(N, 3) * 0.08                       # sum of squares straddles the 0.03 reset threshold
    inputs["rot_err"] = f.rot_err.numpy().copy()
    f.spin_coef, f.hand_reset_step = 1.0, 0
    f.emergence_reward = torch.zeros(N); f.heap_movement_penalty = torch.zeros(N)
    f.dist_reward_scale, f.rot_reward_scale, f.rot_eps, f.action_penalty_scale = -1.0, 1.0, 0.1, -0.0
    f.success_tolerance, f.reach_goal_bonus, f.fall_dist, f.fall_penalty, f.rotation_id = 0.1, 250.0, 0.4, 0.0, 1
    f.max_consecutive_successes, f.av_factor, f.object_type = 0, to_torch(0.1), "egg"
    f.meta_rew_buf = torch.zeros(N); f.extras = {}
    f.total_steps = 0; f.print_success_stat = False
    with torch.no_grad():
        IS.BlockAssemblyInsertSim.compute_reward(f, f.actions)
    np.savez(os.path.join(OUT, "insert_post_physics.npz"), obs=f.obs_buf.numpy(), states=f.states_buf.numpy(), rew=f.rew_buf.numpy(),
             reset=f.reset_buf.numpy(), finger_dist=f.arm_hand_finger_dist.numpy(), extra_target_pos=f.extra_target_pos.numpy(),
             consec=f.consecutive_successes.numpy(), consec_in=np.array([0.7], np.float32), **inputs)
    print("insert: rew range", float(f.rew_buf.min()), float(f.rew_buf.max()), "resets", int(f.reset_buf.sum()), "bonus envs", int((f.rew_buf > 1).sum()))

    # ---- pre_physics_step, no-reset branch (IS:1495-1565)
    p = Fake()
    p.num_envs, p.device = N, "cpu"
    p.gym, p.sim = mock.MagicMock(), None
    p.reset_buf = torch.zeros(N, dtype=torch.long); p.reset_goal_buf = torch.zeros(N, dtype=torch.long)
    p.test_robot_controller = False; p.use_teleoperation = False; p.apply_teleoper_perturbation = False
    p.actuated_dof_indices = torch.arange(7, 23)
    p.arm_hand_dof_lower_limits, p.arm_hand_dof_upper_limits = lo, hi
    p.act_moving_average = 1.0                                                          # yaml:16
    p.prev_targets = lo + (hi - lo) * torch.rand(N, 23)
    p.cur_targets = p.prev_targets.clone()
    p.rigid_body_states = rb
    p.hand_base_rigid_body_index = 7
    p.target_euler = to_torch([0.0, 3.1415, 1.571]).repeat((N, 1))                      # IS:448
    jac = torch.randn(N, 23, 6, 23) * 0.4
    p.jacobian_tensor = jac
    p.arm_hand_dof_pos = dof_state[..., 0].clone()
    acts = torch.rand(N, 23) * 2 - 1
    pin = dict(prev_targets=p.prev_targets.numpy().copy(), hand_pose=rb[:, 7, 0:7].numpy().copy(), jac7=jac[:, 6, :, :7].numpy().copy(),
               dof_pos=p.arm_hand_dof_pos.numpy().copy(), actions=acts.numpy().copy(), hand_target_quat=quat_from_euler_xyz(*p.target_euler[0]).numpy())
    IS.BlockAssemblyInsertSim.pre_physics_step(p, acts)
    np.savez(os.path.join(OUT, "insert_pre_physics.npz"), cur_targets=p.cur_targets.numpy(), rot_err=p.rot_err.numpy(), **pin)

    # ---- reset_idx (IS:1328-1493) on a stand-in with a 6-row grasp bank per brick type
    class Quat:                                     # gymapi.Quat().from_euler_zyx(roll, pitch, yaw)
        def from_euler_zyx(self, a, b, c):
            q = quat_from_euler_zyx(float(a), float(b), float(c))
            o = Quat(); o.x, o.y, o.z, o.w = q
            return o
    IS.gymapi.Quat = Quat
    PER = 6
    r = Fake()
    r.num_envs, r.device = N, "cpu"
    r.gym, r.sim = mock.MagicMock(), None
    r.record_completion_time, r.save_hdf5, r.randomize, r.train_t_value, r.replan = False, False, False, False, False
    r.total_steps = 11
    r.num_arm_hand_dofs = 23
    r.x_unit_tensor = torch.tensor([1.0, 0, 0]).repeat(N, 1)
    r.y_unit_tensor = torch.tensor([0, 1.0, 0]).repeat(N, 1)
    root2 = root.clone()
    r.root_state_tensor = root2
    r.hand_indices, r.object_indices, r.extra_object_indices = f.hand_indices, f.object_indices, f.extra_object_indices
    r.goal_object_indices = f.hand_indices + 2
    r.lego_indices = (f.hand_indices[:, None] + 9 + torch.arange(8)[None]).long()
    r.lego_segmentation_indices = f.lego_segmentation_indices.clone()
    r.pre_exchange_lego_segmentation_indices = f.lego_segmentation_indices.clone()
    r.segmentation_target_rot, r.segmentation_target_pos = root[f.lego_segmentation_indices, 3:7].clone(), root[f.lego_segmentation_indices, 0:3].clone()
    r.extra_target_rot, r.extra_target_pos = f.extra_target_rot.clone(), f.extra_target_pos.clone()
    r.symmetry_extra_target_rot = f.symmetry_extra_target_rot.clone()
    r.success_buf = torch.zeros(N, 2)
    r.rigid_body_states = rb.clone()
    r.base_pos = r.rigid_body_states[:, 0, 0:3]
    r.rb_forces = torch.zeros(N, nb_env, 3)
    r.object_init_state = torch.zeros(N, 13); r.object_init_state[:, 0:3] = torch.tensor([0.0, 0.0, -10.78]); r.object_init_state[:, 6] = 1
    r.goal_states = r.object_init_state.clone(); r.goal_init_state = r.object_init_state.clone()
    r.goal_displacement_tensor = torch.tensor([-0.2, -0.06, 0.12])
    r.reset_goal_buf = torch.zeros(N, dtype=torch.long)
    r.reset_position_noise, r.up_axis_idx = 0.0, 2
    r.object_pose_for_open_loop = torch.zeros(N, 7)
    lego_init = torch.zeros(N, 8, 13)
    for i in range(8):                               # IS:723: parked beside the table on the ground
        lego_init[:, i, 0:3] = torch.tensor([1.13 + 0.13 * (i % 3) + 0.1, -0.23 + 0.23 * (i // 3), 0.02])
        lego_init[:, i, 3:7] = torch.tensor(quat_from_euler_zyx(0.0, 0.0, 0.785))
    r.lego_init_states = lego_init
    r.saved_grasping_object_ternimal_states_list = [torch.cat([torch.randn(PER, 1, 3) * 0.05 + torch.tensor([0.2, -0.1, 0.9]), rq(PER, 1), torch.randn(PER, 1, 6)], -1)
                                                    for _ in range(8)]
    r.saved_grasping_hand_ternimal_states_list = [torch.stack([lo + (hi - lo) * torch.rand(PER, 23), torch.randn(PER, 23)], -1) for _ in range(8)]
    dof2 = dof_state.clone()
    r.dof_state = dof2.view(N * 23, 2)
    r.arm_hand_dof_pos = dof2[..., 0]
    r.prev_targets, r.cur_targets = torch.randn(N, 23), torch.randn(N, 23)
    r.t_value_obs_buf = torch.zeros(N, 7)
    r.random_force_prob = torch.zeros(N); r.force_prob_range = to_torch([0.001, 0.1])
    r.segmentation_target_init_pos, r.segmentation_target_init_rot = torch.zeros(N, 3), torch.zeros(N, 4)
    r.progress_buf = torch.tensor(rng.integers(1, 125, size=N), dtype=torch.long)
    r.reset_buf = torch.zeros(N, dtype=torch.long)
    env_ids = torch.tensor([0, 3, 4, 5, 9, 13, 14, 22])
    r.reset_buf[env_ids] = 1
    r.successes = torch.rand(N); r.meta_rew_buf = torch.rand(N)
    r.reset_target_pose = lambda ids, apply_reset=False: IS.BlockAssemblyInsertSim.reset_target_pose(r, ids, apply_reset)
    rin = dict(root=root2.numpy().copy(), dof_state=dof2.numpy().copy(), env_ids=env_ids.numpy(), progress=r.progress_buf.numpy().copy(),
               successes=r.successes.numpy().copy(), lego_init=lego_init.numpy(),
               bank_obj=torch.stack(r.saved_grasping_object_ternimal_states_list).numpy(), bank_hand=torch.stack(r.saved_grasping_hand_ternimal_states_list).numpy(),
               seg_rot=r.segmentation_target_rot.numpy(), seg_pos=r.segmentation_target_pos.numpy(), extra_rot=r.extra_target_rot.numpy(),
               extra_pos=r.extra_target_pos.numpy())
    # the reference draws random.sample(range(0, 5000), 1) per env (IS:1449-1451): here PER rows are banked, so the draw is patched to
    # a recorded slot sequence (the oracle / kernel draw theirs from Philox; the test feeds the same slots)
    slots = [int(x) for x in rng.integers(0, PER, size=len(env_ids))]
    calls = {"n": 0}
    real_sample = random.sample

    def fake_sample(pop, k):
        if isinstance(pop, range) and len(pop) == 5000:
            s = slots[calls["n"]]; calls["n"] += 1
            return [s]
        return real_sample(pop, k)
    random.seed(5)
    plate_rot = None
    with mock.patch.object(IS.random, "sample", fake_sample), mock.patch.object(IS, "print", lambda *a, **k: None, create=True):
        IS.BlockAssemblyInsertSim.reset_idx(r, env_ids, torch.tensor([], dtype=torch.long))
        plate_rot = r.target_rot_rand[0]
    np.savez(os.path.join(OUT, "insert_reset.npz"), root_out=root2.numpy(), dof_out=dof2.numpy(), prev_targets=r.prev_targets.numpy(),
             cur_targets=r.cur_targets.numpy(), init_pos=r.segmentation_target_init_pos.numpy(), init_rot=r.segmentation_target_init_rot.numpy(),
             progress_out=r.progress_buf.numpy(), reset_out=r.reset_buf.numpy(), successes_out=r.successes.numpy(), success_buf=r.success_buf.numpy(),
             slots=np.asarray(slots), plate_rot=np.int64(plate_rot), **rin)
    print("insert golden vectors written to", os.path.normpath(OUT), "| insertion successes at reset:", r.success_buf[env_ids, 0].tolist())


if __name__ == "__main__":
    main()
