"""ToolPositioningGrasp / ToolPositioningOrient (BASELINE configs[4]): pin the CPU oracle to golden vectors produced by EXECUTING the
reference's own Python (oracle/gen_golden_tool.py: compute_observations twice in a row, compute_hand_reward, pre_physics_step and
reset_idx of tasks/tool_positioning/allegro_hand_tool_positioning_{grasp,orient}.py with Isaac Gym stubbed)."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NA = 11      # actors per env in the generator's stand-in: hand, object, goal, table, 5 bin boxes, the tool, the plate
TASKS = {"grasp": "ToolPositioningGrasp", "orient": "ToolPositioningOrient"}


def _scene(name):
    from seqdex_b200.tasks.cfg import scene_from_cfg
    return scene_from_cfg(TASKS[name])


def _load(name):
    return dict(np.load(os.path.join(G, name)))


def _rows72(tool_rows):
    n = tool_rows.shape[0]
    out = np.zeros((n, 72, 13), np.float32)
    out[..., 6] = 1
    out[:, 0] = tool_rows
    return out


def test_tool_scene_constants():
    for name, (task_id, ep) in {"grasp": (4, 150), "orient": (5, 125)}.items():
        c = _scene(name).c
        assert c.task == task_id and c.max_episode_length == ep and c.n_bricks == 1 and c.n_bshapes == 2          # yaml:6; TG:762 one tool, two boxes
        assert abs(c.max_depen_vel - 1.0) < 1e-7 and abs(c.contact_offset - 0.002) < 1e-9                          # yaml sim block
        d = _load(f"tool_{name}_pre.npz")
        np.testing.assert_allclose(list(c.hand_target_quat), d["hand_target_quat"], atol=1e-7)                    # TG:506, 1628
        r = _load(f"tool_{name}_reset.npz")
        out = r["root_out"].reshape(-1, NA, 13)
        np.testing.assert_allclose(list(c.tool_plate_pose), out[r["env_ids"][0], 10, 0:7], atol=1e-7)              # TG:1505-1512
    g = _load("tool_grasp_reset.npz")
    c = _scene("grasp").c
    out = g["root_out"].reshape(-1, NA, 13)
    np.testing.assert_allclose(list(c.tool_reset_pos), out[g["env_ids"][0], 9, 0:3], atol=1e-7)                    # TG:1496-1498
    np.testing.assert_allclose(list(c.prepare_arm), g["dof_out"][g["env_ids"][0], :7, 0], atol=1e-7)               # TG:284
    # mass of the two-box tool at the URDF's density (567): handle 3.2 x 33 x 3.6 cm + head 3.1 x 4.7 x 13.3 cm
    assert abs(1.0 / c.br_invm[0] - 567.0 * (0.032 * 0.33 * 0.036 + 0.031 * 0.047 * 0.133)) < 1e-6


@pytest.mark.parametrize("name", ["grasp", "orient"])
def test_tool_post_physics_matches_reference(name, oracle_lib):
    d = _load(f"tool_{name}_post.npz")
    n = len(d["progress0"])
    o = oracle_lib.OracleEnv(_scene(name), n)
    assert o.obs.shape == (n, 468) and o.states.shape == (n, 564)                    # TG:224-243, TO:170-189
    o.target_init[:, 0:3] = d["init_pos"]
    o.target_init[:, 3:7] = d["init_rot"]
    o.obs[:, 0:312] = d["hist_obs"].reshape(n, 312)          # the reference's history frames: newest, then one older
    o.states[:, 0:376] = d["hist_states"].reshape(n, 376)
    o.consec[:] = d["consec_in"]
    o.reset[:] = d["reset_in0"]
    o.successes[:] = d["successes_in0"]
    for call in (0, 1):
        root = d[f"root{call}"].reshape(n, NA, 13)
        o.set_brick_roots(_rows72(root[:, 9]))
        o.link[:] = d[f"rb{call}"][:, :24]
        o.dof[:, 0, :23] = d[f"dof_state{call}"][..., 0]
        o.dof[:, 1, :23] = d[f"dof_state{call}"][..., 1]
        o.actions[:] = d[f"actions{call}"]
        o.plate[:] = root[:, 10, 0:7]
        o.progress[:] = d[f"progress{call}"] - 1             # post_physics_step increments first (TG:1678)
        np.testing.assert_array_equal(o.reset, d[f"reset_in{call}"])
        np.testing.assert_array_equal(o.successes, d[f"successes_in{call}"])
        o.post_physics()
        np.testing.assert_allclose(o.obs, d[f"obs{call}"], rtol=0, atol=3e-6)
        np.testing.assert_allclose(o.states, d[f"states{call}"], rtol=0, atol=3e-6)
        assert float(np.abs(o.states[:, 141]).max()) == 0.0                          # never written (TG:1308, TO:1172)
        np.testing.assert_allclose(o.rew, d[f"rew{call}"], rtol=3e-5, atol=2e-7)     # own exp / asin polynomials vs libm
        assert np.array_equal(o.reset, d[f"reset{call}"])
        np.testing.assert_array_equal(o.successes, d[f"successes{call}"])
        np.testing.assert_allclose(o.finger_dist, d[f"finger_dist{call}"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(o.consec, d[f"consec{call}"], rtol=1e-6)
    # the second call's older frames are the first call's newest (TG:1334-1336, 1366-1368)
    np.testing.assert_allclose(o.obs[:, 156:312], d["obs0"][:, 0:156], rtol=0, atol=3e-6)
    np.testing.assert_allclose(o.states[:, 188:376], d["states0"][:, 0:188], rtol=0, atol=3e-6)
    if name == "grasp":
        assert (d["rew0"] > 1).sum() >= 4 and 4 <= d["reset1"].sum() < n and d["successes1"].sum() >= 1, \
            "golden set must exercise the orientation bonus, the move-out / time-out resets and the successes flag"
    else:
        assert (d["rew0"] > 1).sum() >= 1 and 2 <= d["reset1"].sum() < n


@pytest.mark.parametrize("name", ["grasp", "orient"])
def test_tool_pre_physics_matches_reference(name, oracle_lib):
    d = _load(f"tool_{name}_pre.npz")
    n = d["actions"].shape[0]
    o = oracle_lib.OracleEnv(_scene(name), n)
    o.dof[:, 0, :23] = d["dof_pos"]
    o.dof[:, 2, :23] = d["prev_targets"]
    o.link[:, 7, 0:7] = d["hand_pose"]
    o.jac7[:] = d["jac7"]
    o.progress[:] = d["progress"]
    o.reset[:] = 0
    o.pre_physics(d["actions"])
    np.testing.assert_allclose(o.dof[:, 2, 7:23], d["cur_targets"][:, 7:], rtol=0, atol=2e-6)
    if name == "orient":
        np.testing.assert_allclose(o.dof[:, 2, :7], d["cur_targets"][:, :7], rtol=0, atol=0)       # the arm holds its (clamped) previous target
        assert d["prev_targets"][0, 0] > d["cur_targets"][0, 0]
    else:
        np.testing.assert_allclose(o.dof[:, 2, :7], d["cur_targets"][:, :7], rtol=2e-3, atol=2e-4)  # LU inverse (torch) vs Cholesky solve, as for control_ik
        parked = d["progress"] > 90
        assert parked.sum() >= 2
        np.testing.assert_allclose(o.dof[parked, 2, :7], d["cur_targets"][parked, :7], rtol=0, atol=1e-7)      # TG:1640
        np.testing.assert_array_equal(o.dof[parked, 2, 7:23], np.clip(d["prev_targets"][parked, 7:], _scene(name).dof_lo[7:], _scene(name).dof_hi[7:]))


def test_tool_grasp_reset_idx_matches_reference(oracle_lib):
    d = _load("tool_grasp_reset.npz")
    n = d["root"].reshape(-1, NA, 13).shape[0]
    o = oracle_lib.OracleEnv(_scene("grasp"), n)
    root = d["root"].reshape(n, NA, 13)
    o.set_brick_roots(_rows72(root[:, 9]))
    o.dof[:, 0, :23] = d["dof_state"][..., 0]
    o.dof[:, 1, :23] = d["dof_state"][..., 1]
    o.plate[:] = root[:, 10, 0:7]
    o.finger_dist[:] = d["finger_dist"]
    o.progress[:] = d["progress"]
    o.successes[:] = d["successes"]
    o.obs[:] = 1.5; o.states[:] = -2.5
    o.reset[:] = 0
    o.reset[d["env_ids"]] = 1
    o.total_steps = 11
    o.gb_index[:] = d["index_in"]
    o.pitch_k, o.yaw_u = int(d["pitch_k"]), d["yaw_u"]
    o._tool_reset_idx()
    out = d["root_out"].reshape(n, NA, 13)
    ids = d["env_ids"]
    rest = np.setdiff1d(np.arange(n), ids)
    got = o.brick_roots()[:, 0]
    np.testing.assert_allclose(got[ids], out[ids, 9], rtol=0, atol=2e-6)               # (0.29, 0.19, 0.675), pitch k x 1.571, yaw u x 3.14, at rest
    np.testing.assert_allclose(got[rest, 0:7], out[rest, 9, 0:7], rtol=0, atol=2e-6)
    np.testing.assert_allclose(o.plate[ids], out[ids, 10, 0:7], rtol=0, atol=1e-7)
    np.testing.assert_array_equal(o.plate[rest], root[rest, 10, 0:7])
    np.testing.assert_allclose(o.dof[ids, 0, :23], d["dof_out"][ids, :, 0], rtol=0, atol=1e-7)
    assert float(np.abs(o.dof[ids, 1, :23]).max()) == 0.0 and float(np.abs(d["dof_out"][ids, :, 1]).max()) == 0.0
    np.testing.assert_allclose(o.dof[ids, 2, :23], d["cur_targets"][ids], rtol=0, atol=1e-7)
    np.testing.assert_allclose(o.dof[ids, 2, :23], d["prev_targets"][ids], rtol=0, atol=1e-7)
    np.testing.assert_array_equal(o.dof[rest, 0, :23], d["dof_state"][rest, :, 0])
    np.testing.assert_allclose(o.target_init[ids, 0:3], d["init_pos"][ids], atol=1e-7)
    np.testing.assert_allclose(o.target_init[ids, 3:7], d["init_rot"][ids], atol=2e-6)
    assert np.array_equal(o.progress, d["progress_out"]) and np.array_equal(o.reset, d["reset_out"])
    np.testing.assert_array_equal(o.successes, d["successes_out"])
    np.testing.assert_array_equal(o.success_buf[ids, 0], d["success_buf"][ids, 0])
    assert 0 < d["success_buf"][ids, 0].sum() < len(ids)
    # history: the reference zeroes obs_buf and both stacks of frames of the resetting envs (TG:1563-1568)
    assert float(np.abs(d["obs_out"][ids]).max()) == 0.0 and float(np.abs(d["frames_obs"][ids]).max()) == 0.0 and float(np.abs(d["frames_states"][ids]).max()) == 0.0
    assert float(np.abs(o.obs[ids]).max()) == 0.0 and float(np.abs(o.states[ids]).max()) == 0.0
    assert np.all(o.obs[rest] == 1.5) and np.all(o.states[rest] == -2.5)
    # banking (TG:1436-1457): same slots, same rows, same ring indices -- including the ring that wraps after slot 10000
    np.testing.assert_array_equal(o.gb_index, d["index_out"])
    where = d["bank_where"]
    assert len(where) >= 4 and (where[:, 1] == 10000).any()
    mask = np.zeros((8, 11024), bool)
    mask[where[:, 0], where[:, 1]] = True
    assert float(np.abs(o.gb_obj[~mask]).max()) == 0.0 and float(np.abs(o.gb_hand[~mask]).max()) == 0.0
    np.testing.assert_array_equal(o.gb_hand[where[:, 0], where[:, 1]], d["bank_hand_rows"])
    np.testing.assert_allclose(o.gb_obj[where[:, 0], where[:, 1]], d["bank_obj_rows"], rtol=0, atol=2e-6)     # root -> COM -> root round trip


def test_tool_orient_reset_idx_matches_reference(oracle_lib):
    d = _load("tool_orient_reset.npz")
    n = d["root"].reshape(-1, NA, 13).shape[0]
    o = oracle_lib.OracleEnv(_scene("orient"), n)
    root = d["root"].reshape(n, NA, 13)
    o.set_brick_roots(_rows72(root[:, 9]))
    o.dof[:, 0, :23] = d["dof_state"][..., 0]
    o.dof[:, 1, :23] = d["dof_state"][..., 1]
    o.plate[:] = root[:, 10, 0:7]
    o.progress[:] = d["progress"]
    o.successes[:] = d["successes"]
    o.obs[:] = 1.5; o.states[:] = -2.5
    o.reset[:] = 0
    o.reset[d["env_ids"]] = 1
    o.total_steps = 11
    o.set_grasp_bank(d["bank_hand"], d["bank_obj"])
    o.slot_by_env = d["slot_by_env"]
    o._tool_reset_idx()
    out = d["root_out"].reshape(n, NA, 13)
    ids = d["env_ids"]
    rest = np.setdiff1d(np.arange(n), ids)
    got = o.brick_roots()[:, 0]
    np.testing.assert_allclose(got[ids], out[ids, 9], rtol=0, atol=3e-6)               # the banked root row, velocities included (TO:1397)
    assert float(np.abs(out[ids, 9, 7:13]).max()) > 0.1
    np.testing.assert_allclose(got[rest], out[rest, 9], rtol=0, atol=3e-6)
    np.testing.assert_allclose(o.plate[ids], out[ids, 10, 0:7], rtol=0, atol=1e-7)
    np.testing.assert_array_equal(o.dof[ids, 0, :23], d["dof_out"][ids, :, 0])
    np.testing.assert_array_equal(o.dof[ids, 1, :23], d["dof_out"][ids, :, 1])         # DoF velocities restored too (TO:1398)
    np.testing.assert_array_equal(o.dof[ids, 2, :23], d["cur_targets"][ids])
    np.testing.assert_allclose(o.target_init[ids, 0:3], d["init_pos"][ids], atol=0)
    np.testing.assert_allclose(o.target_init[ids, 3:7], d["init_rot"][ids], atol=0)
    np.testing.assert_allclose(o.target_init[ids], d["t_value_obs"][ids], atol=0)      # t_value_obs_buf = the restored pose (TO:1400)
    assert np.array_equal(o.progress, d["progress_out"]) and np.array_equal(o.reset, d["reset_out"])
    np.testing.assert_array_equal(o.successes, d["successes_out"])
    np.testing.assert_array_equal(o.success_buf[ids, 0], d["success_buf"][ids, 0])
    assert 0 < d["success_buf"][ids, 0].sum() < len(ids)
    assert np.all(o.obs == 1.5) and np.all(o.states == -2.5)                            # TO keeps its history across resets


def test_tool_orient_online_tvalue_labels_match_reference(oracle_lib):
    """TO:1305-1316, executed by the reference with `if_t_value` switched on: success_buf for ALL envs = within 1 cm of the plate and within
    0.1 rad of its orientation or the pi-about-z twin"""
    d = _load("tool_orient_reset.npz")
    n = d["tv_obs_in"].shape[0]
    o = oracle_lib.OracleEnv(_scene("orient"), n)
    rows = np.zeros((n, 13), np.float32)
    rows[:, 0:3], rows[:, 3:7] = d["tv_target_pos"], d["tv_target_rot"]
    o.set_brick_roots(_rows72(rows))
    o.plate[:] = d["tv_plate"]
    label = o.tool_tvalue_labels()
    np.testing.assert_array_equal(o.success_buf, d["tv_success_buf"])
    np.testing.assert_array_equal(label, (d["tv_success_buf"][:, 0] < 0.5).astype(np.int32))
    assert 4 <= d["tv_success_buf"][:, 0].sum() < n and d["tv_success_buf"][2, 0] == 0.0      # env 2: aligned but 2 cm away


def test_tool_chain_insertion_observations_match_reference(oracle_lib):
    """ToolPositioningChain.compute_insertion_observations (TC:1404-1440) executed by the reference right after the second
    compute_observations call of the Grasp golden: the inner policy's actions in 23:46, the inner clock in slot 60, its own history"""
    d = _load("tool_grasp_post.npz")
    n = len(d["progress0"])
    o = oracle_lib.OracleEnv(_scene("grasp"), n)
    o.obs[:] = d["obs1"]
    ins = np.zeros((n, 468), np.float32)
    ins[:, 0:312] = d["ins_hist"].reshape(n, 312)
    o.tool_insertion_obs(d["ins_actions"], d["ins_progress"], ins)
    np.testing.assert_array_equal(ins, d["ins_obs"])
    np.testing.assert_array_equal(ins[:, 23:46], d["ins_actions"])
    np.testing.assert_array_equal(ins[:, 156:468], d["ins_hist"].reshape(n, 312))
