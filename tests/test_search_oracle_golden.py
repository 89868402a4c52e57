"""BlockAssemblySearch (BASELINE configs[0]): pin the CPU oracle to golden vectors produced by EXECUTING the reference's own
Python (oracle/gen_golden_search.py: compute_observations, compute_hand_reward, pre_physics_step of
tasks/block_assembly/allegro_hand_block_assembly_search.py with Isaac Gym stubbed), then run the task the way the reference's
own CPU-runnable configuration does: num_envs = 4."""
import os

import numpy as np
import pytest

from seqdex_b200.camera import SEARCH_CAMERA, look_at

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sscene():
    from seqdex_b200.tasks.cfg import scene_from_cfg
    return scene_from_cfg("BlockAssemblySearch")   # the yaml-stated sim / env parameters (contact_offset 0.02)


def _load(name):
    return dict(np.load(os.path.join(G, name)))


def test_search_scene_constants(sscene, scene):
    c = sscene.c
    assert c.task == 2 and c.max_episode_length == 75 and abs(c.act_moving_average - 0.6) < 1e-7        # yaml:6,16
    init = np.ctypeslib.as_array(c.brick_init).reshape(72, 13)
    init_gs = np.ctypeslib.as_array(scene.c.brick_init).reshape(72, 13)
    np.testing.assert_allclose(init[:, 2] - init_gs[:, 2], 0.06, atol=1e-6)                             # SE:637 0.68 vs GS:737 0.62
    assert abs(np.ctypeslib.as_array(c.fixed_root).reshape(60, 13)[0, 2] - 0.63) < 1e-6                 # SE:661
    d = _load("search_pre_physics.npz")
    np.testing.assert_allclose(list(c.hand_target_quat), d["hand_target_quat"], atol=1e-7)              # SE:1565-1566 in fp32
    assert abs(c.default_dof[0] - 0.9467) < 1e-6 and abs(c.prepare_dof[1] + 0.49826458) < 1e-6          # SE:208, 220


def test_mask_features_match_the_reference_reduction(oracle_lib):
    """pixels / int(mean row) / int(mean column) of the reference's own loop over real (synthetic) images (SE:1231-1241)"""
    d = _load("search_post_physics.npz")
    out = np.zeros(3, np.int32)
    for e, m in enumerate(d["masks"]):
        m = np.ascontiguousarray(m, np.uint8)
        oracle_lib.lib().sdxo_mask_features(m.ctypes.data_as(__import__("ctypes").c_void_p), 128, 128, oracle_lib.ip(out))
        assert out.tolist() == d["seg"][e].tolist(), (e, out, d["seg"][e])
    assert (d["seg"][:, 0] == 0).sum() >= 3 and (d["seg"][:, 0] > 100).sum() >= 5


def test_search_post_physics_matches_reference(sscene, oracle_lib):
    d = _load("search_post_physics.npz")
    n = len(d["progress"])
    o = oracle_lib.OracleEnv(sscene, n)
    root = d["root"].reshape(n, 142, 13)
    o.set_brick_roots(np.ascontiguousarray(root[:, 9:81]))
    o.link[:] = d["rb"][:, :24]
    o.netf[:] = d["contact"]
    o.dof[:, 0, :23] = d["dof_state"][..., 0]
    o.dof[:, 1, :23] = d["dof_state"][..., 1]
    o.actions[:] = d["actions"]
    o.target_init[:, 0:3] = d["init_pos"]; o.target_init[:, 3:7] = d["init_rot"]
    o.progress[:] = d["progress"] - 1
    o.progress[0] = 10                                     # keep env 0's clock away from the end-of-episode camera branch
    o.reset[:] = d["reset_in"]
    o.obs[:] = d["prev_obs"]; o.states[:] = d["prev_states"]; o.tvobs[:] = d["prev_tvobs"]
    o.successes[:] = d["successes"]; o.consec[:] = d["consec_in"]
    o.seg[:] = d["seg"]
    o.post_physics()
    ok = np.arange(n) != 0                                 # env 0 ran with another progress value (see above)
    np.testing.assert_allclose(o.obs, d["obs"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(o.states, d["states"], rtol=0, atol=3e-6)
    assert np.array_equal(o.states[:, 188:], d["prev_states"][:, 188:])          # Search never shifts a state history (SE:1168-1218)
    np.testing.assert_allclose(o.tvobs, d["tvobs"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(o.rew, d["rew"], rtol=2e-6, atol=2e-4)             # terms of size 100: 1000 x clamp(...)
    assert np.array_equal(o.reset[ok], d["reset"][ok])
    np.testing.assert_allclose(o.consec, d["consec"], rtol=1e-6) if d["reset"][0] == 0 else None
    assert d["reset"].sum() >= 3 and d["rew"].min() < -50 and d["rew"].max() > 20


def test_search_pre_physics_matches_reference(sscene, oracle_lib):
    d = _load("search_pre_physics.npz")
    n = len(d["dof_pos"])
    o = oracle_lib.OracleEnv(sscene, n)
    o.dof[:, 0, :23] = d["dof_pos"]
    o.dof[:, 2, :23] = d["prev_targets"]
    o.link[:, 7, 0:7] = d["hand_pose"]
    o.jac7[:] = d["jac7"]
    o.reset[:] = 0
    rows = o.brick_roots()
    for e in range(n):
        rows[e, sscene.target_brick_index(e), 0:3] = d["target_pos"][e]
    o.set_brick_roots(rows)
    o.pre_physics(d["actions"])
    np.testing.assert_allclose(o.dof[:, 2, :23], d["cur_targets"], rtol=2e-3, atol=5e-4)          # IK conditioning (LU vs Cholesky)
    np.testing.assert_allclose(o.dof[:, 2, 7:23], d["cur_targets"][:, 7:23], rtol=0, atol=1e-6)   # finger EMA: exact arithmetic


def test_search_runs_as_the_reference_cpu_configuration(sscene, oracle_lib):
    """BASELINE configs[0]: BlockAssemblySearch, num_envs = 4.  One whole episode and the reset after it: 60 settle steps per
    reset, the end-of-episode render with the hand parked (emergence reward), banking of heaps whose target shows enough pixels."""
    n = 4
    o = oracle_lib.OracleEnv(sscene, n)
    o.set_camera(look_at(**SEARCH_CAMERA))
    o.enable_search_bank(3)
    rng = np.random.default_rng(3)
    o.step(rng.uniform(-1, 1, size=(n, 23)).astype(np.float32))
    assert o.last_reset_sim_steps == 60 and (o.progress == 1).all() and (o.episode == 1).all()
    tz = o.brick_roots()[np.arange(n), [sscene.target_brick_index(e) for e in range(n)], 2]
    assert (tz < 0.9).all() and (tz > 0.6).all()               # dropped from 0.9 m into the bin (SE:1394)
    np.testing.assert_allclose(o.dof[:, 0, :7], np.broadcast_to(np.ctypeslib.as_array(sscene.c.prepare_dof)[:7], (n, 7)), atol=0.2)
    pix0 = o.seg[:, 0].copy()
    for t in range(73):
        o.step(rng.uniform(-1, 1, size=(n, 23)).astype(np.float32))
        assert o.last_reset_sim_steps == 0
        if t < 72:
            assert np.array_equal(o.seg[:, 0], pix0)           # between renders the features are those of the last image
    assert o.reset.all() and (o.progress == 74).all()
    np.testing.assert_allclose(o.emergence, (o.seg[:, 0] - pix0) * 5.0)          # SE:1645
    assert np.abs(o.tvobs[:, -65:-3]).sum() > 0 and np.allclose(o.tvobs[:, -1], o.seg[:, 0] / 100.0)
    expect_banked = [int(o.seg[e, 0] > [20, 20, 15, 20][e]) for e in range(n)]
    o.step(rng.uniform(-1, 1, size=(n, 23)).astype(np.float32))
    assert o.last_reset_sim_steps == 60 and (o.progress == 1).all() and (o.episode == 2).all()
    assert o.sb_index[:4].tolist() == expect_banked and o.sb_index[4:].sum() == 0
    for e in range(n):
        if expect_banked[e]:
            assert np.abs(o.sb_rows[e, 0]).sum() > 0 and np.abs(o.sb_hand[e, 0]).sum() > 0
