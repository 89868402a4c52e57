"""BlockAssemblyOrient: pin the CPU oracle to golden vectors produced by EXECUTING the reference's own Python
(oracle/gen_golden_orient.py: compute_observations, compute_hand_reward, pre_physics_step of
tasks/block_assembly/allegro_hand_block_assembly_orient.py with Isaac Gym stubbed), then exercise the scripted reset."""
import os

import numpy as np
import pytest

from tests.util import lattice_bank

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def oscene():
    from seqdex_b200.tasks.cfg import scene_from_cfg
    return scene_from_cfg("BlockAssemblyOrient")   # the yaml-stated sim / env parameters (contact_offset 0.02)


def _load(name):
    return dict(np.load(os.path.join(G, name)))


def test_orient_scene_constants(oscene, scene):
    c = oscene.c
    assert c.task == 1 and c.max_episode_length == 75 and abs(c.act_moving_average - 0.2) < 1e-7     # yaml:6,16
    assert list(c.dof_kp)[7:] == [20.0] * 16 and abs(c.dof_effort[7] - 0.7) < 1e-6                   # OR:588-598
    assert list(scene.c.dof_kp)[7:] == [50.0] * 16                                                   # GraspSim unchanged
    d = _load("orient_pre_physics.npz")
    np.testing.assert_allclose(list(c.hand_target_quat), d["hand_target_quat"], atol=1e-7)           # OR:484,1738 in fp32


def test_orient_post_physics_matches_reference(oscene, oracle_lib):
    d = _load("orient_post_physics.npz")
    n = len(d["progress"])
    o = oracle_lib.OracleEnv(oscene, n, tvalue_weights=d["tv_weights"])
    assert o.obs.shape == (n, 186) and o.states.shape == (n, 564)                                    # OR:206-208
    root = d["root"].reshape(n, 142, 13)
    o.set_brick_roots(np.ascontiguousarray(root[:, 9:81]))
    o.link[:] = d["rb"][:, :24]
    o.dof[:, 0, :23] = d["dof_state"][..., 0]
    o.dof[:, 1, :23] = d["dof_state"][..., 1]
    o.actions[:] = d["actions"]
    o.target_init[:, 0:3] = d["init_pos"]
    o.target_init[:, 3:7] = d["init_rot"]
    o.progress[:] = d["progress"] - 1          # post_physics_step increments first (OR:1781)
    o.reset[:] = d["reset_in"]
    o.obs[:] = d["prev_obs"]
    o.states[:] = d["prev_states"]
    o.successes[:] = d["successes"]
    o.consec[:] = d["consec_in"]
    o.post_physics()
    # compute_real_observations writes 48 of the 62 slots of frame 0 and nothing else (OR:1308-1326): the rest keeps its value
    np.testing.assert_allclose(o.obs, d["obs"], rtol=0, atol=3e-6)
    assert np.array_equal(o.obs[:, 16:30], d["prev_obs"][:, 16:30]) and np.array_equal(o.obs[:, 62:], d["prev_obs"][:, 62:])
    np.testing.assert_allclose(o.states, d["states"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(o.rew, d["rew"], rtol=2e-5, atol=1e-7)        # exp(-(5 z + 5 d)): own exp polynomial vs libm
    assert np.array_equal(o.reset, d["reset"])
    clear = np.abs(d["tvalue_raw"] - 0.99) > 1e-4                             # gate threshold: skip ulp-borderline cases
    assert clear.sum() >= 20 and np.array_equal(o.tvalue[clear], d["tvalue"][clear])
    assert 4 <= d["tvalue"].sum() <= 20, "golden set must exercise both gate outcomes"
    np.testing.assert_allclose(o.finger_dist, d["finger_dist"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(o.consec, d["consec"], rtol=1e-6)
    assert d["reset"].sum() >= 3 and d["rew"].max() > 0.9 and d["rew"].min() < 1e-3


def test_orient_pre_physics_matches_reference(oscene, oracle_lib):
    d = _load("orient_pre_physics.npz")
    n = len(d["progress"])
    o = oracle_lib.OracleEnv(oscene, n)
    o.dof[:, 0, :23] = d["dof_pos"]
    o.dof[:, 2, :23] = d["prev_targets"]
    o.link[:, 7, 0:7] = d["hand_pose"]
    o.jac7[:] = d["jac7"]
    o.progress[:] = d["progress"]
    o.target_init[:, 0:3] = d["init_pos"]
    o.reset[:] = 0
    # put each env's target brick where the golden root tensor has it (root rows -> COM-frame brick block)
    rows = o.brick_roots()
    for e in range(n):
        rows[e, oscene.target_brick_index(e), 0:3] = d["target_pos"][e]
    o.set_brick_roots(rows)
    o.pre_physics(d["actions"])
    np.testing.assert_allclose(o.dof[:, 2, :23], d["cur_targets"], rtol=2e-3, atol=5e-4)   # IK solve conditioning (LU vs Cholesky)
    np.testing.assert_allclose(o.dof[:, 2, 7:23], d["cur_targets"][:, 7:23], rtol=0, atol=1e-6)   # finger EMA: exact arithmetic
    assert (d["progress"] > 75).sum() >= 2, "golden set must exercise the progress > 75 branch (OR:1735, 1743)"


def test_orient_scripted_reset(oscene, oracle_lib):
    """reset_idx / post_reset (OR:1390-1695): 0 extra contact steps when nobody resets, 53 on the very first reset
    (total_steps == 0 skips the lift + banking), 103 afterwards; the hand ends above the target's initial pose; heaps whose
    brick passes the gate are banked in env order with the reference's wrap-around."""
    n = 16
    o = oracle_lib.OracleEnv(oscene, n)
    o.set_heap_bank(lattice_bank(oscene, 3))
    o.enable_orient_heap_bank(4)
    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, size=(n, 23)).astype(np.float32)
    o.step(a)                                   # BT:63 reset_buf starts at 1 -> first reset, no banking
    assert o.last_reset_sim_steps == 53 and (o.progress == 1).all() and not o.reset.any()
    assert (o.episode == 1).all() and o.ob_index.sum() == 0
    # target_init is the target brick's pose after the 2 settle steps of post_reset (OR:1617-1624), not the banked row
    assert np.abs(o.target_init[:, 2] - 0.62).max() < 0.05
    hb = o.link[:, 7, 0:3]
    # approach script: hand 22 cm above / 18 cm behind the initial target pose (OR:1666-1672), IK converged to a few cm
    want = o.target_init[:, 0:3] + np.array([-0.18, 0.0, 0.22], np.float32)
    assert np.abs(hb - want).max() < 0.06, np.abs(hb - want).max()
    o.step(a)
    assert o.last_reset_sim_steps == 0 and (o.progress == 2).all()
    # force the gate open, let the episode time out: lockstep reset with lift + banking
    o.tv[-1] += 50.0                            # output bias of the 'feasible' logit
    o.progress[:] = 73
    o.step(a)
    assert o.reset.all()                        # progress 74 >= 75 - 1 (OR:1866-1867)
    tg_y = o.brick_roots()[np.arange(n), [oscene.target_brick_index(e) for e in range(n)], 1]
    o.step(a)
    assert o.last_reset_sim_steps == 103 and (o.progress == 1).all() and (o.episode == 2).all()
    assert o.ob_index.sum() > 0                 # the lattice heaps have the target brick at 0 < y < 0.5: banked
    # 16 envs, 8 types -> 2 envs per type, ring wrap 4: index 2 per type that banked twice
    assert set(o.ob_index.tolist()) <= {0, 1, 2}
    banked = o.ob_rows[0, 0]
    assert np.isfinite(banked).all() and np.abs(banked[:, 3:7]).sum() > 0
    assert tg_y.min() > 0.0 and tg_y.max() < 0.5
