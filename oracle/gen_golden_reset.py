#!/usr/bin/env python
"""Golden vectors for reset_idx of BlockAssemblyGraspSim, produced by EXECUTING THE REFERENCE'S OWN PYTHON
(`tasks/block_assembly/allegro_hand_block_assembly_grasp_sim.py:1361-1553`, with reset_target_pose `:1334-1358`) on a stand-in
`self` (stubs as in gen_golden.py; Isaac Gym calls are mocks).  Runs only in the build container; the output is committed as
tests/golden/reset_idx.npz.

The reference draws the heap it restores with `random.sample(range(0, 5000), 1)` per env (`:1507-1510`); the oracle draws it from
its own Philox stream.  To compare the two, `random.sample` is patched to return, for env e, the slot the
oracle's stream selects (seed, env, episode) -- everything else that runs is the reference's code: the grasp terminal-state
banking gate and ring bookkeeping (`:1398-1445`), the root-state writes of the restored heap (`:1507-1513`), the hand reset
(`:1524-1536`), segmentation_target_init_* (`:1547-1548`) and the per-env counters (`:1550-1553`).
"""
import os
import random
import sys
from unittest import mock

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import install_stubs, Fake, OUT   # noqa: E402

N, PER_TYPE, NA = 16, 4, 142                      # envs, heaps per brick type in the stand-in pickle, actors per env
SEED = 22


def main():
    install_stubs()
    import tasks.block_assembly.allegro_hand_block_assembly_grasp_sim as GS
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from seqdex_b200.scene import Scene, quat_from_euler_zyx
    from oracle import dr_oracle
    scene = Scene()

    class Quat:                                     # gymapi.Quat().from_euler_zyx (the written pose is dead: the heap row replaces it)
        def from_euler_zyx(self, a, b, c):
            q = quat_from_euler_zyx(float(a), float(b), float(c))
            o = Quat(); o.x, o.y, o.z, o.w = q
            return o
    GS.gymapi.Quat = Quat

    torch.manual_seed(4321)
    rng = np.random.default_rng(4321)

    def rq(*shape):
        q = torch.randn(*shape, 4)
        return q / q.norm(dim=-1, keepdim=True)

    f = Fake()
    f.num_envs, f.device = N, "cpu"
    f.gym, f.sim = mock.MagicMock(), None
    f.record_completion_time, f.save_hdf5, f.randomize = False, False, False
    f.total_steps = 7
    f.num_arm_hand_dofs = 23
    f.z_unit_tensor = torch.tensor([0, 0, 1.0]).repeat(N, 1)
    f.x_unit_tensor = torch.tensor([1.0, 0, 0]).repeat(N, 1)
    f.y_unit_tensor = torch.tensor([0, 1.0, 0]).repeat(N, 1)
    # ---- actors: 0 hand, 1 object, 2 goal, 3-8 table + bin, 9-140 the 132 bricks (72 free + 60 fixed), 141 base plate
    f.hand_indices = torch.arange(N) * NA
    f.object_indices = f.hand_indices + 1
    f.goal_object_indices = f.hand_indices + 2
    f.lego_indices = (f.hand_indices.view(N, 1) + 9 + torch.arange(132).view(1, 132))
    seg = torch.tensor([Scene.target_brick_index(e) for e in range(N)])
    f.lego_segmentation_indices = f.hand_indices + 9 + seg
    root = torch.zeros(N * NA, 13)
    root[:, 0:3] = torch.randn(N * NA, 3) * 0.1 + torch.tensor([0.25, 0.0, 0.7])
    root[:, 3:7] = rq(N * NA)
    root[:, 7:13] = torch.randn(N * NA, 6) * 0.2
    fixed = torch.from_numpy(np.asarray(scene.fixed_root, np.float32))
    for e in range(N):
        root[e * NA + 9 + 72:e * NA + 9 + 132] = fixed
    # the banking gate (GS:1402-1404): target brick's y < 0, finger distance < 0.6, t-value > 0.8 -- every combination occurs
    ysign = torch.tensor([-1.0 if (e % 4) != 3 else 1.0 for e in range(N)])
    root[f.lego_segmentation_indices, 1] = ysign * (0.05 + 0.1 * torch.rand(N))
    f.root_state_tensor = root
    f.segmentation_target_pos = root[f.lego_segmentation_indices, 0:3].clone()
    f.segmentation_target_rot = root[f.lego_segmentation_indices, 3:7].clone()
    f.arm_hand_finger_dist = torch.tensor([0.3 if (e % 5) != 4 else 0.9 for e in range(N)])
    f.tvalue = torch.tensor([0.95 if (e % 3) != 2 else 0.5 for e in range(N)])
    f.reset_buf = torch.tensor([1 if (e % 8) != 6 else 0 for e in range(N)], dtype=torch.long)
    env_ids = f.reset_buf.nonzero(as_tuple=False).squeeze(-1)
    f.dof_state = torch.randn(N * 23, 2) * 0.4
    f.arm_hand_dof_state = f.dof_state.view(N, -1, 2)[:, :23]
    f.arm_hand_dof_pos = f.arm_hand_dof_state[..., 0]
    f.arm_hand_dof_vel = f.arm_hand_dof_state[..., 1]
    f.arm_hand_dof_default_vel = torch.zeros(23)
    lo = torch.from_numpy(np.asarray(scene.dof_lo[:23], np.float32)); hi = torch.from_numpy(np.asarray(scene.dof_hi[:23], np.float32))
    f.arm_hand_dof_lower_limits, f.arm_hand_dof_upper_limits = lo, hi
    prepare = torch.from_numpy(np.concatenate([np.ctypeslib.as_array(scene.c.prepare_arm), np.zeros(16)]).astype(np.float32))
    f.arm_hand_prepare_dof_pos_list = [prepare.clone()]
    f.arm_hand_prepare_dof_poses = torch.zeros(N, 23)
    f.end_effector_rot_list = [torch.tensor([0, 0, 0, 1.0])]
    f.end_effector_rotation = torch.zeros(N, 4)
    f.prev_targets, f.cur_targets = torch.randn(N, 23), torch.randn(N, 23)
    # ---- the grasp terminal-state rings (GS:397-407) with a ring that is about to wrap (index 5000 -> > 5000 -> 0, GS:1441-1443)
    f.saved_grasp_hand_ternimal_states_list = [torch.zeros(5008, 23, 2) for _ in range(8)]
    f.saved_grasp_object_ternimal_states_list = [torch.zeros(5008, 13) for _ in range(8)]
    f.saved_grasp_ternimal_states_index_list = [0, 3, 5000, 17, 0, 4999, 1, 2]
    index_before = list(f.saved_grasp_ternimal_states_index_list)
    f.can_save = [0] * 8
    # ---- the stand-in for saved_searching_ternimal_states_good_mo_tvalue.pkl (GS:412-413): list[8] of [PER_TYPE, 132, 13]
    bank = torch.zeros(8, PER_TYPE, 132, 13)
    bank[..., 0:3] = torch.randn(8, PER_TYPE, 132, 3) * 0.08 + torch.tensor([0.25, -0.05, 0.68])
    bank[..., 3:7] = rq(8, PER_TYPE, 132)
    bank[..., 7:13] = torch.randn(8, PER_TYPE, 132, 6) * 0.3          # velocities in the pickle: zeroed by GS:1513
    bank[:, :, 72:] = fixed
    f.saved_searching_ternimal_states_list = [bank[t].clone() for t in range(8)]
    episode = rng.integers(0, 50, size=N).astype(np.int32)
    slots = [int(dr_oracle.philox4x32(SEED, np.array([e], np.uint64), int(episode[e]), 1)[0][0]) % PER_TYPE for e in range(N)]
    # ---- the rest reset_idx touches
    f.perturb_steps, f.perturb_direction = torch.zeros(N), torch.zeros(N, 6)
    f.rigid_body_states = torch.randn(N, 165, 13)
    f.base_pos = torch.zeros(N, 3)
    f.rb_forces = torch.zeros(N, 165, 3)
    f.object_init_state = torch.zeros(N, 13); f.object_init_state[:, 2] = -10.78; f.object_init_state[:, 6] = 1
    f.reset_position_noise, f.up_axis_idx = 0.0, 2
    f.object_pose_for_open_loop = torch.zeros(N, 7)
    f.lego_init_states = root[f.lego_indices.view(-1)].clone().view(N, 132, 13)
    f.goal_states = torch.zeros(N, 13); f.goal_init_state = torch.zeros(N, 13); f.goal_displacement_tensor = torch.zeros(3)
    f.reset_goal_buf = torch.zeros(N, dtype=torch.long)
    f.random_force_prob = torch.zeros(N); f.force_prob_range = torch.tensor([0.001, 0.1])
    f.segmentation_target_init_pos, f.segmentation_target_init_rot = torch.zeros(N, 3), torch.zeros(N, 4)
    f.progress_buf = torch.arange(N, dtype=torch.long) + 3
    f.successes = torch.ones(N)
    f.meta_rew_buf = torch.ones(N)
    f.reset_target_pose = lambda ids, apply_reset=False: GS.BlockAssemblyGraspSim.reset_target_pose(f, ids, apply_reset)

    inputs = dict(root_before=root.clone().numpy().reshape(N, NA, 13), dof_before=f.dof_state.clone().numpy().reshape(N, 23, 2),
                  targets_before=f.cur_targets.clone().numpy(), bank=bank[:, :, :72].numpy().copy(), reset=f.reset_buf.numpy().copy(),
                  finger_dist=f.arm_hand_finger_dist.numpy().copy(), tvalue=f.tvalue.numpy().copy(), episode=episode,
                  slots=np.array(slots, np.int32), index_before=np.array(index_before, np.int32), seed=np.array([SEED], np.int64),
                  progress_before=f.progress_buf.numpy().copy())
    real_sample = random.sample
    it = iter([slots[int(e)] for e in env_ids])

    def sample(pop, k):
        if len(pop) == 5000:                       # GS:1509: the heap to restore
            return [next(it)]
        return real_sample(pop, k)
    with mock.patch.object(random, "sample", sample):
        GS.BlockAssemblyGraspSim.reset_idx(f, env_ids, env_ids)
    idx_after = np.array(f.saved_grasp_ternimal_states_index_list, np.int32)
    out = dict(root_after=f.root_state_tensor.numpy().reshape(N, NA, 13)[:, 9:81].copy(),
               dof_after=f.dof_state.numpy().reshape(N, 23, 2).copy(), prev_targets=f.prev_targets.numpy().copy(),
               cur_targets=f.cur_targets.numpy().copy(), progress=f.progress_buf.numpy().copy(), reset_after=f.reset_buf.numpy().copy(),
               successes=f.successes.numpy().copy(), target_init_pos=f.segmentation_target_init_pos.numpy().copy(),
               target_init_rot=f.segmentation_target_init_rot.numpy().copy(), index_after=idx_after,
               gb_hand=np.stack([t[:5008].numpy() for t in f.saved_grasp_hand_ternimal_states_list]),
               gb_obj=np.stack([t[:5008].numpy() for t in f.saved_grasp_object_ternimal_states_list]))
    # keep the file small: only the ring slots that can have been written
    touched = sorted({int(i) for i in index_before} | {int(i) + 1 for i in index_before} | {0, 1, 2})
    out["gb_slots"] = np.array(touched, np.int32)
    out["gb_hand"] = out["gb_hand"][:, touched]
    out["gb_obj"] = out["gb_obj"][:, touched]
    np.savez_compressed(os.path.join(OUT, "reset_idx.npz"), **inputs, **out)
    print("wrote reset_idx.npz; envs reset:", env_ids.tolist(), "ring index", index_before, "->", idx_after.tolist())


if __name__ == "__main__":
    main()
