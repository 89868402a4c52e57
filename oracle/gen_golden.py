#!/usr/bin/env python
"""Generate golden vectors for the oracle by EXECUTING THE REFERENCE'S OWN PYTHON.

Runs only in the build container (it imports /root/reference, which does not exist on the GPU box);
the outputs are committed under tests/golden/ and are what `-m "not gpu"` tests pin the oracle to.

The reference task module cannot be imported as is: it needs Isaac Gym (closed binary), matplotlib, cv2,
pytorch3d ....  We install stub modules for those, with ONE piece of real
content: ``isaacgym.torch_utils``, restated below from NVIDIA's published IsaacGymEnvs
``isaacgymenvs/utils/torch_jit_utils.py`` (quats xyzw).  Everything else that
runs is the reference's code, unmodified:
    compute_hand_reward       GS:1706-1776  (TorchScript, called directly)
    control_ik                GS:1796-1804
    compute_observations      GS:1090-1218  (+ compute_sim_observations GS:1299-1332,
                                              compute_contact_asymmetric_observations GS:1220-1280)
    pre_physics_step          GS:1555-1638  (no-reset branch)
    GraspInsertTValue         TVF:30-46
called as unbound methods on a stand-in ``self`` that carries exactly the attributes they read.
"""
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

REF = "/root/reference/dexteroushandenvs"
OUT = os.environ.get("SEQDEX_GOLDEN_OUT") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")   # override: tests/test_golden_regen.py


# ---------------------------------------------------------------- isaacgym.torch_utils (public restatement)
def _torch_utils():
    m = types.ModuleType("isaacgym.torch_utils")

    def to_torch(x, dtype=torch.float, device='cpu', requires_grad=False):
        return torch.tensor(x, dtype=dtype, device=device, requires_grad=requires_grad)

    @torch.jit.script
    def quat_mul(a, b):
        assert a.shape == b.shape
        shape = a.shape
        a = a.reshape(-1, 4)
        b = b.reshape(-1, 4)
        x1, y1, z1, w1 = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
        x2, y2, z2, w2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
        ww = (z1 + x1) * (x2 + y2)
        yy = (w1 - y1) * (w2 + z2)
        zz = (w1 + y1) * (w2 - z2)
        xx = ww + yy + zz
        qq = 0.5 * (xx + (z1 - x1) * (x2 - y2))
        w = qq - ww + (z1 - y1) * (y2 - z2)
        x = qq - xx + (x1 + w1) * (x2 + w2)
        y = qq - yy + (w1 - x1) * (y2 + z2)
        z = qq - zz + (z1 + y1) * (w2 - x2)
        return torch.stack([x, y, z, w], dim=-1).view(shape)

    @torch.jit.script
    def normalize(x, eps: float = 1e-9):
        return x / x.norm(p=2, dim=-1).clamp(min=eps, max=None).unsqueeze(-1)

    @torch.jit.script
    def quat_apply(a, b):
        shape = b.shape
        a = a.reshape(-1, 4)
        b = b.reshape(-1, 3)
        xyz = a[:, :3]
        t = xyz.cross(b, dim=-1) * 2
        return (b + a[:, 3:] * t + xyz.cross(t, dim=-1)).view(shape)

    @torch.jit.script
    def quat_conjugate(a):
        shape = a.shape
        a = a.reshape(-1, 4)
        return torch.cat((-a[:, :3], a[:, -1:]), dim=-1).view(shape)

    @torch.jit.script
    def quat_from_angle_axis(angle, axis):
        theta = (angle / 2).unsqueeze(-1)
        xyz = normalize(axis) * theta.sin()
        w = theta.cos()
        return normalize(torch.cat([xyz, w], dim=-1))

    @torch.jit.script
    def tf_inverse(q, t):
        q_inv = quat_conjugate(q)
        return q_inv, -quat_apply(q_inv, t)

    @torch.jit.script
    def tf_combine(q1, t1, q2, t2):
        return quat_mul(q1, q2), quat_apply(q1, t2) + t1

    @torch.jit.script
    def scale(x, lower, upper):
        return (0.5 * (x + 1.0) * (upper - lower) + lower)

    @torch.jit.script
    def unscale(x, lower, upper):
        return (2.0 * x - upper - lower) / (upper - lower)

    @torch.jit.script
    def tensor_clamp(t, min_t, max_t):
        return torch.max(torch.min(t, max_t), min_t)

    def torch_rand_float(lower, upper, shape, device):
        return (upper - lower) * torch.rand(*shape, device=device) + lower

    for k, v in dict(locals()).items():
        if k != "m":
            setattr(m, k, v)
    m.__all__ = [k for k in dict(locals()) if k not in ("m", "k", "v")]
    return m


def install_stubs():
    ig = types.ModuleType("isaacgym")
    ig.gymapi = mock.MagicMock(name="gymapi")
    ig.gymtorch = mock.MagicMock(name="gymtorch")
    ig.gymutil = mock.MagicMock(name="gymutil")
    ig.gymtorch.unwrap_tensor = lambda t: t
    ig.torch_utils = _torch_utils()
    sys.modules.update({"isaacgym": ig, "isaacgym.gymapi": ig.gymapi, "isaacgym.gymtorch": ig.gymtorch,
                        "isaacgym.gymutil": ig.gymutil, "isaacgym.torch_utils": ig.torch_utils})
    for name in ("matplotlib", "matplotlib.pyplot", "PIL", "PIL.Image", "cv2", "pyquaternion", "pytorch3d",
                 "pytorch3d.transforms", "gym", "gym.spaces", "h5py", "torchvision"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mock.MagicMock(name=name)
    sys.path.insert(0, REF)


class Fake:
    pass


def main():
    os.makedirs(OUT, exist_ok=True)
    install_stubs()
    import tasks.block_assembly.allegro_hand_block_assembly_grasp_sim as GS
    from policy_sequencing.terminal_value_function import GraspInsertTValue
    from isaacgym.torch_utils import to_torch

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from seqdex_b200.scene import Scene, quat_from_euler_zyx
    scene = Scene()
    torch.manual_seed(1234)
    rng = np.random.default_rng(1234)
    N = 24

    def rq(*shape):
        q = torch.randn(*shape, 4)
        return q / q.norm(dim=-1, keepdim=True)

    # ---- 1. GraspInsertTValue forward + sigmoid (TVF:30-46, GS:1200-1201)
    tv = GraspInsertTValue(input_dim=4, output_dim=2)
    wts = torch.cat([p.detach().reshape(-1) for p in (tv.linear1.weight, tv.linear1.bias, tv.linear2.weight, tv.linear2.bias,
                                                      tv.linear3.weight, tv.linear3.bias, tv.output_layer.weight,
                                                      tv.output_layer.bias)]).numpy().astype(np.float32)
    qin = rq(64)
    with torch.no_grad():
        tv_out = torch.sigmoid(tv(qin))[:, 1]
    np.savez(os.path.join(OUT, "tvalue.npz"), weights=wts, qin=qin.numpy(), out=tv_out.numpy())

    # ---- 2. control_ik (GS:1796-1804)
    J = torch.randn(N, 6, 7) * 0.5
    dpose = torch.randn(N, 6, 1) * 0.3
    u = GS.control_ik(J, "cpu", dpose, N)
    np.savez(os.path.join(OUT, "control_ik.npz"), J=J.numpy(), dpose=dpose.squeeze(-1).numpy(), u=u.numpy())

    # ---- 3. compute_observations + compute_reward on a stand-in self
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
    # make half of the envs "close" so the reward's lift branch and the 0.6 m reset branch are both exercised
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
    f.contact_tensor = torch.randn(N, nb_env * 3) * 0.2
    f.sensor_handle_indices = torch.tensor([1, 2, 3, 4, 5, 6])
    f.envs = [None]
    f.camera_offset_quat = to_torch(quat_from_euler_zyx(0.0, -3.141 + 0.5, 1.571))
    f.camera_offset_pos = to_torch([0.03, 0.107 - 0.098, 0.067 + 0.107])
    f.segmentation_target_init_pos = root[f.lego_segmentation_indices, 0:3] + torch.randn(N, 3) * 0.05
    f.segmentation_target_init_rot = rq(N)
    f.actions = torch.rand(N, 23) * 2 - 1
    f.perturb_direction = torch.zeros(N, 6)
    f.progress_buf = torch.tensor(rng.integers(0, 150, size=N), dtype=torch.long)
    f.progress_buf[0] = 148; f.progress_buf[1] = 149; f.progress_buf[2] = 74; f.progress_buf[3] = 75
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
    f.one_frame_num_obs, f.one_frame_num_states = 132, 188
    prev_obs = torch.randn(N, 396) * 0.3
    prev_states = torch.randn(N, 564) * 0.3
    f.obs_buf, f.states_buf = prev_obs.clone(), prev_states.clone()
    f.obs_buf_stack_frames = [prev_obs[:, 0:132].clone(), prev_obs[:, 132:264].clone(), torch.zeros(N, 132)]
    f.state_buf_stack_frames = [prev_states[:, 0:188].clone(), prev_states[:, 188:376].clone(), torch.zeros(N, 188)]
    f.compute_sim_observations = lambda *a, **k: GS.BlockAssemblyGraspSim.compute_sim_observations(f, *a, **k)
    f.compute_contact_asymmetric_observations = lambda: GS.BlockAssemblyGraspSim.compute_contact_asymmetric_observations(f)
    inputs = dict(rb=rb.numpy().copy(), root=root.numpy().copy(), dof_state=dof_state.numpy().copy(), actions=f.actions.numpy().copy(),
                  init_pos=f.segmentation_target_init_pos.numpy().copy(), init_rot=f.segmentation_target_init_rot.numpy().copy(),
                  prev_obs=prev_obs.numpy(), prev_states=prev_states.numpy(), progress=f.progress_buf.numpy().copy(),
                  seg_index=seg.numpy(), tv_weights=wts)
    with torch.no_grad():
        GS.BlockAssemblyGraspSim.compute_observations(f)
    # compute_reward (GS:1060-1067)
    f.rew_buf = torch.zeros(N)
    f.reset_buf = torch.zeros(N, dtype=torch.long); f.reset_buf[5] = 1
    inputs["reset_in"] = f.reset_buf.numpy().copy()
    f.reset_goal_buf = torch.zeros(N, dtype=torch.long)
    f.successes = torch.zeros(N); f.successes[5] = 2.0
    inputs["successes"] = f.successes.numpy().copy()
    f.consecutive_successes = torch.tensor([0.7])
    f.spin_coef, f.hand_reset_step, f.max_episode_length = 1.0, 0, 150
    f.object_pos, f.object_rot, f.object_angvel = root[f.object_indices, 0:3], root[f.object_indices, 3:7], root[f.object_indices, 10:13]
    f.emergence_reward = torch.zeros(N); f.heap_movement_penalty = torch.zeros(N)
    f.dist_reward_scale, f.rot_reward_scale, f.rot_eps, f.action_penalty_scale = -1.0, 1.0, 0.1, -0.0
    f.success_tolerance, f.reach_goal_bonus, f.fall_dist, f.fall_penalty, f.rotation_id = 0.1, 250.0, 0.4, 0.0, 1
    f.max_consecutive_successes, f.av_factor, f.object_type = 0, to_torch(0.1), "egg"
    f.meta_rew_buf = torch.zeros(N); f.extras = {}
    f.total_steps = 0; f.print_success_stat = False
    with torch.no_grad():
        GS.BlockAssemblyGraspSim.compute_reward(f, f.actions)
    np.savez(os.path.join(OUT, "post_physics.npz"), obs=f.obs_buf.numpy(), states=f.states_buf.numpy(), rew=f.rew_buf.numpy(),
             reset=f.reset_buf.numpy(), tvalue=f.tvalue.detach().numpy(), finger_dist=f.arm_hand_finger_dist.numpy(),
             consec=f.consecutive_successes.numpy(), consec_in=np.array([0.7], np.float32), **inputs)

    # ---- 4. pre_physics_step, no-reset branch (GS:1570-1638)
    p = Fake()
    p.num_envs, p.device = N, "cpu"
    p.gym, p.sim = mock.MagicMock(), None
    p.reset_buf = torch.zeros(N, dtype=torch.long); p.reset_goal_buf = torch.zeros(N, dtype=torch.long)
    p.test_robot_controller = False; p.use_teleoperation = False; p.apply_teleoper_perturbation = False
    p.actuated_dof_indices = torch.arange(7, 23)
    p.arm_hand_dof_lower_limits, p.arm_hand_dof_upper_limits = lo, hi
    p.act_moving_average = 1.0
    p.prev_targets = lo + (hi - lo) * torch.rand(N, 23)
    p.cur_targets = p.prev_targets.clone()
    p.progress_buf = torch.tensor(rng.integers(0, 150, size=N), dtype=torch.long)
    p.progress_buf[0:6] = torch.tensor([75, 76, 100, 101, 125, 126])
    p.segmentation_target_init_pos = torch.randn(N, 3) * 0.1 + torch.tensor([0.2, 0.2, 0.65])
    p.rigid_body_states = rb
    p.hand_base_rigid_body_index = 7
    jac = torch.randn(N, 23, 6, 23) * 0.4
    p.jacobian_tensor = jac
    p.arm_hand_dof_pos = dof_state[..., 0].clone()
    p.arm_hand_insertion_prepare_dof_pos_list = [to_torch([-0.1560, -0.2140, -0.2795, -2.1806, -0.0681, 1.9730, 1.1735]),
                                                 to_torch([-0.1800, -0.1604, -0.2770, -2.2674, -0.0533, 2.1049, 1.1696])]
    acts = torch.rand(N, 23) * 2 - 1
    pin = dict(prev_targets=p.prev_targets.numpy().copy(), progress=p.progress_buf.numpy().copy(),
               init_pos=p.segmentation_target_init_pos.numpy().copy(), hand_pos=rb[:, 7, 0:3].numpy().copy(),
               jac7=jac[:, 6, :, :7].numpy().copy(), dof_pos=p.arm_hand_dof_pos.numpy().copy(), actions=acts.numpy().copy())
    GS.BlockAssemblyGraspSim.pre_physics_step(p, acts)
    np.savez(os.path.join(OUT, "pre_physics.npz"), cur_targets=p.cur_targets.numpy(), **pin)
    print("golden vectors written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
