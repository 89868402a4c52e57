"""INTEGRATION.md section 2 as a test: the REFERENCE's own ``compute_observations`` and ``compute_reward`` (tasks/block_assembly/
allegro_hand_block_assembly_grasp_sim.py, executed unmodified with Isaac Gym stubbed as in oracle/gen_golden.py) run on the
Isaac-Gym-shaped tensors the facade materialised on a B200 (``sdx_refresh`` -> tests/golden/facade_dump.npz, tools/dump_facade.py) and
arrive at what the fused kernels computed from the same state.  Needs /root/reference (build container only)."""
import os
import sys
from unittest import mock

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/dexteroushandenvs"), reason="the reference tree is not present on this box")


def test_reference_task_python_on_the_facade_tensors_matches_the_kernels():
    import torch
    sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
    from gen_golden import Fake, install_stubs
    install_stubs()
    import tasks.block_assembly.allegro_hand_block_assembly_grasp_sim as GS
    from isaacgym.torch_utils import to_torch
    from policy_sequencing.terminal_value_function import GraspInsertTValue
    from seqdex_b200.scene import Scene, quat_from_euler_zyx
    d = dict(np.load(os.path.join(G, "facade_dump.npz")))
    n = d["rb"].shape[0]
    scene = Scene()
    tv = GraspInsertTValue(input_dim=4, output_dim=2)
    off = 0
    with torch.no_grad():
        for p in (tv.linear1.weight, tv.linear1.bias, tv.linear2.weight, tv.linear2.bias, tv.linear3.weight, tv.linear3.bias, tv.output_layer.weight,
                  tv.output_layer.bias):
            p.copy_(torch.from_numpy(d["tv_weights"][off:off + p.numel()]).view_as(p)); off += p.numel()
    f = Fake()
    f.num_envs, f.device = n, "cpu"
    f.gym, f.sim = mock.MagicMock(), None
    # ---- the facade's tensors, in Isaac Gym's layouts and with the task's own index bookkeeping (GS:907-1000)
    f.root_state_tensor = torch.from_numpy(d["root"])                      # [n * 142, 13]
    f.rigid_body_states = torch.from_numpy(d["rb"])                        # [n, 165, 13]
    dof_state = torch.from_numpy(d["dof_state"])                           # [n, 23, 2]
    f.arm_hand_dof_pos, f.arm_hand_dof_vel = dof_state[..., 0], dof_state[..., 1]
    contact = torch.zeros(n, 165, 3)
    contact[:, :24] = torch.from_numpy(d["netf"])                          # net contact forces exist for the robot's links
    f.contact_tensor = contact.view(n, -1)
    f.hand_indices = torch.arange(n) * 142
    f.object_indices = f.hand_indices + 1
    f.extra_object_indices = f.hand_indices + 141
    f.lego_segmentation_indices = f.hand_indices + 9 + torch.tensor([Scene.target_brick_index(e) for e in range(n)])
    f.goal_states = torch.zeros(n, 13)
    f.hand_base_rigid_body_index = f.mount_rigid_body_index = 7
    f.fingertip_handles = torch.tensor([11, 19, 23, 15])
    f.sensor_handle_indices = torch.tensor([1, 2, 3, 4, 5, 6])
    f.envs = [None]
    f.camera_offset_quat = to_torch(quat_from_euler_zyx(0.0, -3.141 + 0.5, 1.571))
    f.camera_offset_pos = to_torch([0.03, 0.107 - 0.098, 0.067 + 0.107])
    # ---- task state the env keeps between steps
    f.segmentation_target_init_pos = torch.from_numpy(d["target_init"][:, 0:3].copy())
    f.segmentation_target_init_rot = torch.from_numpy(d["target_init"][:, 3:7].copy())
    f.actions = torch.from_numpy(d["actions"])
    f.progress_buf = torch.from_numpy(d["progress"])
    f.perturb_direction, f.perturb_steps = torch.zeros(n, 6), torch.zeros(n, 1)
    f.z_unit_tensor, f.x_unit_tensor = to_torch([0, 0, 1]).repeat(n, 1), to_torch([1, 0, 0]).repeat(n, 1)
    f.t_value, f.obs_type, f.save_hdf5 = tv, "partial_contact", False
    f.arm_hand_dof_lower_limits, f.arm_hand_dof_upper_limits = torch.from_numpy(scene.dof_lo), torch.from_numpy(scene.dof_hi)
    f.vel_obs_scale, f.one_frame_num_obs, f.one_frame_num_states = 0.2, 132, 188
    po, ps = torch.from_numpy(d["prev_obs"]), torch.from_numpy(d["prev_states"])
    f.obs_buf, f.states_buf = po.clone(), ps.clone()
    f.obs_buf_stack_frames = [po[:, 0:132].clone(), po[:, 132:264].clone(), torch.zeros(n, 132)]
    f.state_buf_stack_frames = [ps[:, 0:188].clone(), ps[:, 188:376].clone(), torch.zeros(n, 188)]
    f.compute_sim_observations = lambda *a, **k: GS.BlockAssemblyGraspSim.compute_sim_observations(f, *a, **k)
    f.compute_contact_asymmetric_observations = lambda: GS.BlockAssemblyGraspSim.compute_contact_asymmetric_observations(f)
    with torch.no_grad():
        GS.BlockAssemblyGraspSim.compute_observations(f)
    np.testing.assert_allclose(f.obs_buf.numpy(), d["obs"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(f.states_buf.numpy(), d["states"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(f.tvalue.detach().numpy(), d["tvalue"], rtol=0, atol=2e-6)
    # ---- compute_reward (GS:1060-1067) on what compute_observations left on `self`
    f.rew_buf = torch.zeros(n)
    f.reset_buf, f.reset_goal_buf = torch.zeros(n, dtype=torch.long), torch.zeros(n, dtype=torch.long)
    f.successes, f.consecutive_successes = torch.zeros(n), torch.tensor([0.0])
    f.spin_coef, f.hand_reset_step, f.max_episode_length = 1.0, 0, 150
    root = f.root_state_tensor
    f.object_pos, f.object_rot, f.object_angvel = root[f.object_indices, 0:3], root[f.object_indices, 3:7], root[f.object_indices, 10:13]
    f.emergence_reward, f.heap_movement_penalty = torch.zeros(n), torch.zeros(n)
    f.dist_reward_scale, f.rot_reward_scale, f.rot_eps, f.action_penalty_scale = -1.0, 1.0, 0.1, -0.0
    f.success_tolerance, f.reach_goal_bonus, f.fall_dist, f.fall_penalty, f.rotation_id = 0.1, 250.0, 0.4, 0.0, 1
    f.max_consecutive_successes, f.av_factor, f.object_type = 0, to_torch(0.1), "egg"
    f.meta_rew_buf, f.extras, f.total_steps, f.print_success_stat = torch.zeros(n), {}, 0, False
    with torch.no_grad():
        GS.BlockAssemblyGraspSim.compute_reward(f, f.actions)
    np.testing.assert_allclose(f.rew_buf.numpy(), d["rew"], rtol=2e-5, atol=2e-7)
    np.testing.assert_array_equal(f.reset_buf.numpy(), d["reset"])
    assert d["reset"][0] == 1 and d["reset"].sum() < n                     # the env whose clock was set to 148 timed out, the others did not
