"""BlockAssemblyInsertSim: pin the CPU oracle to golden vectors produced by EXECUTING the reference's own Python
(oracle/gen_golden_insert.py: compute_observations, compute_hand_reward, pre_physics_step and reset_idx of
tasks/block_assembly/allegro_hand_block_assembly_insert_sim.py with Isaac Gym stubbed)."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NA = 18      # actors per env in the generator's stand-in: hand, object, goal, table, 5 bin boxes, 8 bricks, base-plate


@pytest.fixture(scope="module")
def iscene():
    from seqdex_b200.tasks.cfg import scene_from_cfg
    return scene_from_cfg("BlockAssemblyInsertSim")


def _load(name):
    return dict(np.load(os.path.join(G, name)))


def _bricks72(rows8):
    n = rows8.shape[0]
    out = np.zeros((n, 72, 13), np.float32)
    out[..., 6] = 1
    out[:, :8] = rows8
    return out


def test_insert_scene_constants(iscene):
    c = iscene.c
    assert c.task == 3 and c.max_episode_length == 125 and abs(c.act_moving_average - 1.0) < 1e-7 and c.n_bricks == 8     # yaml:6,16; IS:689-736
    mods = [(c.st_mod[i], c.st_rem[i]) for i in range(c.n_static)]
    assert mods[-3:] == [(3, 0), (3, 1), (3, 2)] and all(m == (0, 0) for m in mods[:-3])                                   # IS:971-977
    d = _load("insert_pre_physics.npz")
    np.testing.assert_allclose(list(c.hand_target_quat), d["hand_target_quat"], atol=1e-7)                                # IS:448, 1530


def test_insert_post_physics_matches_reference(iscene, oracle_lib):
    d = _load("insert_post_physics.npz")
    n = len(d["progress"])
    o = oracle_lib.OracleEnv(iscene, n)
    assert o.obs.shape == (n, 75) and o.states.shape == (n, 188)                     # IS:172-193
    root = d["root"].reshape(n, NA, 13)
    o.set_brick_roots(_bricks72(root[:, 9:17]))
    o.link[:] = d["rb"][:, :24]
    o.dof[:, 0, :23] = d["dof_state"][..., 0]
    o.dof[:, 1, :23] = d["dof_state"][..., 1]
    o.actions[:] = d["actions"]
    o.target_init[:, 0:3] = d["init_pos"]
    o.target_init[:, 3:7] = d["init_rot"]
    o.plate[:] = root[:, 17, 0:7]
    o.rot_err[:] = d["rot_err"]
    o.progress[:] = d["progress"] - 1          # post_physics_step increments first (IS:1568)
    o.reset[:] = d["reset_in"]
    o.obs[:] = d["prev_obs"]
    o.states[:] = d["prev_states"]
    o.successes[:] = d["successes"]
    o.consec[:] = d["consec_in"]
    o.post_physics()
    np.testing.assert_allclose(o.obs, d["obs"], rtol=0, atol=3e-6)
    assert np.array_equal(o.obs[:, 16:23], d["prev_obs"][:, 16:23]) and np.array_equal(o.obs[:, 60], d["prev_obs"][:, 60])   # never written (IS:1280-1298)
    np.testing.assert_allclose(o.states, d["states"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(o.states[:, 181:184], d["extra_target_pos"], atol=1e-6)
    np.testing.assert_allclose(o.rew, d["rew"], rtol=3e-5, atol=2e-7)          # exp(-rot - 20 d) with own exp / asin polynomials vs libm
    assert np.array_equal(o.reset, d["reset"])
    np.testing.assert_allclose(o.finger_dist, d["finger_dist"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(o.consec, d["consec"], rtol=1e-6)
    assert (d["rew"] > 1).sum() >= 4 and 6 <= d["reset"].sum() < n, "golden set must exercise the success bonus and all three reset causes"


def test_insert_pre_physics_matches_reference(iscene, oracle_lib):
    d = _load("insert_pre_physics.npz")
    n = d["actions"].shape[0]
    o = oracle_lib.OracleEnv(iscene, n)
    o.dof[:, 0, :23] = d["dof_pos"]
    o.dof[:, 2, :23] = d["prev_targets"]
    o.link[:, 7, 0:7] = d["hand_pose"]
    o.jac7[:] = d["jac7"]
    o.reset[:] = 0
    o.pre_physics(d["actions"])
    np.testing.assert_allclose(o.rot_err, d["rot_err"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(o.dof[:, 2, 7:23], d["cur_targets"][:, 7:], rtol=0, atol=2e-6)
    np.testing.assert_allclose(o.dof[:, 2, :7], d["cur_targets"][:, :7], rtol=2e-3, atol=2e-4)     # LU inverse (torch) vs Cholesky solve, as for control_ik


def test_insert_reset_idx_matches_reference(iscene, oracle_lib):
    d = _load("insert_reset.npz")
    n = d["root"].reshape(-1, NA, 13).shape[0]
    o = oracle_lib.OracleEnv(iscene, n)
    root = d["root"].reshape(n, NA, 13)
    o.set_brick_roots(_bricks72(root[:, 9:17]))
    # the episode ended with the target brick where seg_pos / seg_rot say (the reference reads the cached observation tensors, IS:1342-1348)
    rows = o.brick_roots()
    from seqdex_b200.scene import Scene
    for e in range(n):
        rows[e, Scene.target_brick_index(e), 0:3] = d["seg_pos"][e]
        rows[e, Scene.target_brick_index(e), 3:7] = d["seg_rot"][e]
    o.set_brick_roots(rows)
    o.dof[:, 0, :23] = d["dof_state"][..., 0]
    o.dof[:, 1, :23] = d["dof_state"][..., 1]
    o.plate[:] = root[:, 17, 0:7]
    o.progress[:] = d["progress"]
    o.successes[:] = d["successes"]
    o.reset[:] = 0
    o.reset[d["env_ids"]] = 1
    o.total_steps = 11
    o.set_grasp_bank(d["bank_hand"], d["bank_obj"])
    o.plate_yaw = int(d["plate_rot"])
    o._insert_reset_idx(slots=d["slots"])
    out = d["root_out"].reshape(n, NA, 13)
    ids = d["env_ids"]
    rest = np.setdiff1d(np.arange(n), ids)
    got = o.brick_roots()[:, :8]
    np.testing.assert_allclose(got[ids], out[ids, 9:17], rtol=0, atol=2e-6)          # bricks parked, target brick from the bank, velocities zero
    np.testing.assert_allclose(got[rest, :, 0:7], out[rest, 9:17, 0:7], rtol=0, atol=2e-6)   # untouched envs untouched
    np.testing.assert_allclose(o.plate[ids], out[ids, 17, 0:7], rtol=0, atol=1e-7)   # (0.25, -0.2, 0.618), yaw index x 1.57
    np.testing.assert_array_equal(o.plate[rest], root[rest, 17, 0:7])
    np.testing.assert_allclose(o.dof[ids, 0, :23], d["dof_out"][ids, :, 0], rtol=0, atol=0)
    assert float(np.abs(o.dof[ids, 1, :23]).max()) == 0.0 and float(np.abs(d["dof_out"][ids, :, 1]).max()) == 0.0
    np.testing.assert_array_equal(o.dof[ids, 2, :23], d["cur_targets"][ids])          # targets = restored positions (IS:1480-1481)
    np.testing.assert_allclose(o.target_init[ids, 0:3], d["init_pos"][ids], atol=0)
    np.testing.assert_allclose(o.target_init[ids, 3:7], d["init_rot"][ids], atol=0)
    assert np.array_equal(o.progress, d["progress_out"]) and np.array_equal(o.reset, d["reset_out"])
    np.testing.assert_array_equal(o.successes, d["successes_out"])
    np.testing.assert_array_equal(o.success_buf[ids], d["success_buf"][ids])
    assert 0 < d["success_buf"][ids, 0].sum() < len(ids)
