"""BlockAssemblySearch on the GPU (csrc/sdx_task_search.cuh + sdx_camera.cuh through the C-ABI) against the CPU oracle:
BIT-EXACT on the golden inputs of the reference's own Python and over whole episodes -- reset from the drop lattice (60 contact
steps + render), the end-of-episode render with the hand parked, the emergence reward, heap banking -- at BASELINE configs[0]'s
num_envs = 4 and at a size that exercises the 8 brick types."""
import os

import numpy as np
import pytest
import torch

from seqdex_b200.camera import SEARCH_CAMERA, look_at

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sscene():
    from seqdex_b200.tasks.cfg import scene_from_cfg
    return scene_from_cfg("BlockAssemblySearch")   # the yaml-stated sim / env parameters (contact_offset 0.02)


def _cmp(name, a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        raise AssertionError(f"{name}: {len(bad)} of {a.size} differ; max abs diff {d.max():.3e} first at {bad[0]} "
                             f"gpu={a[tuple(bad[0])]!r} oracle={b[tuple(bad[0])]!r}")


def _pair(sscene, oracle_lib, n):
    from seqdex_b200.env import SdxEnv
    g, o = SdxEnv(sscene, n), oracle_lib.OracleEnv(sscene, n)
    cam = look_at(**SEARCH_CAMERA)
    g.set_camera(cam); o.set_camera(cam)
    return g, o


def _all(g, o, tag):
    torch.cuda.synchronize()
    for name, ov in (("OBS", o.obs), ("STATES", o.states), ("TVOBS", o.tvobs), ("REW", o.rew), ("RESET", o.reset), ("PROGRESS", o.progress),
                     ("SEG", o.seg), ("EMERGENCE", o.emergence), ("BRICK", o.brick), ("DOF", o.dof), ("LINK", o.link), ("JAC7", o.jac7),
                     ("TARGET_INIT", o.target_init), ("SLEEP", o.slp), ("EPISODE", o.episode), ("CONSEC", o.consec)):
        _cmp(f"{tag}: {name}", g.tensor(name), ov)


def test_search_kernels_on_golden_inputs(sscene, oracle_lib):
    d = dict(np.load(os.path.join(G, "search_post_physics.npz")))
    n = len(d["progress"])
    g, o = _pair(sscene, oracle_lib, n)
    assert tuple(g.tensor("OBS").shape) == (n, 186) and tuple(g.tensor("TVOBS").shape) == (n, 650)
    root = d["root"].reshape(n, 142, 13)
    o.set_brick_roots(np.ascontiguousarray(root[:, 9:81]))
    o.link[:] = d["rb"][:, :24]; o.netf[:] = d["contact"]
    o.dof[:, 0, :23] = d["dof_state"][..., 0]; o.dof[:, 1, :23] = d["dof_state"][..., 1]
    o.actions[:] = d["actions"]
    o.target_init[:, 0:3] = d["init_pos"]; o.target_init[:, 3:7] = d["init_rot"]
    o.progress[:] = d["progress"] - 1; o.progress[0] = 10
    o.reset[:] = d["reset_in"]
    o.obs[:] = d["prev_obs"]; o.states[:] = d["prev_states"]; o.tvobs[:] = d["prev_tvobs"]
    o.successes[:] = d["successes"]; o.consec[:] = d["consec_in"]
    o.seg[:] = d["seg"]
    for name, v in (("BRICK", o.brick), ("LINK", o.link), ("NETF", o.netf), ("DOF", o.dof), ("ACTIONS", o.actions), ("TARGET_INIT", o.target_init),
                    ("PROGRESS", o.progress), ("RESET", o.reset), ("OBS", o.obs), ("STATES", o.states), ("TVOBS", o.tvobs),
                    ("SUCCESSES", o.successes), ("CONSEC", o.consec), ("SEG", o.seg)):
        g.tensor(name).copy_(torch.from_numpy(v))
    g.post_physics(); o.post_physics()
    torch.cuda.synchronize()
    for name, ov in (("OBS", o.obs), ("STATES", o.states), ("TVOBS", o.tvobs), ("REW", o.rew), ("RESET", o.reset), ("PROGRESS", o.progress),
                     ("CONSEC", o.consec)):
        _cmp(name, g.tensor(name), ov)
    np.testing.assert_allclose(g.tensor("STATES").cpu().numpy(), d["states"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(g.tensor("TVOBS").cpu().numpy(), d["tvobs"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(g.tensor("REW").cpu().numpy(), d["rew"], rtol=2e-6, atol=2e-4)
    p = dict(np.load(os.path.join(G, "search_pre_physics.npz")))
    g, o = _pair(sscene, oracle_lib, n)
    o.dof[:, 0, :23] = p["dof_pos"]; o.dof[:, 2, :23] = p["prev_targets"]
    o.link[:, 7, 0:7] = p["hand_pose"]; o.jac7[:] = p["jac7"]
    o.reset[:] = 0
    rows = o.brick_roots()
    for e in range(n):
        rows[e, sscene.target_brick_index(e), 0:3] = p["target_pos"][e]
    o.set_brick_roots(rows)
    for name, v in (("BRICK", o.brick), ("LINK", o.link), ("DOF", o.dof), ("JAC7", o.jac7), ("RESET", o.reset)):
        g.tensor(name).copy_(torch.from_numpy(v))
    a = p["actions"] * 1.3
    g.pre_physics(torch.from_numpy(a).cuda()); o.pre_physics(a)
    torch.cuda.synchronize()
    _cmp("targets", g.tensor("DOF"), o.dof)
    _cmp("actions", g.tensor("ACTIONS"), o.actions)


@pytest.mark.parametrize("n", [4, 24])
def test_search_episode_bit_exact(sscene, oracle_lib, n):
    """configs[0] (num_envs = 4) and a 24-env run: first reset, a whole episode, the end-of-episode render, the reset with
    banking, and a few steps of the next episode"""
    g, o = _pair(sscene, oracle_lib, n)
    g.enable_search_bank(2); o.enable_search_bank(2)
    rng = np.random.default_rng(n)

    def step(tag, check=True):
        a = rng.uniform(-1, 1, size=(n, 23)).astype(np.float32)
        g.step(torch.from_numpy(a).cuda()); o.step(a)
        assert g.last_reset_sim_steps() == o.last_reset_sim_steps, tag
        if check:
            _all(g, o, tag)

    step("first reset")
    assert o.last_reset_sim_steps == 60
    for t in range(73):
        step(f"step {t}", check=t in (0, 1, 30, 71, 72))
    assert o.reset.all() and (o.emergence != 0).any()
    step("second reset")
    assert o.last_reset_sim_steps == 60
    rows, hand, index = g.search_bank()
    _cmp("bank index", index, o.sb_index)
    _cmp("bank rows", rows, o.sb_rows)
    _cmp("bank hand", hand, o.sb_hand)
    assert o.sb_index.sum() > 0
    for t in range(3):
        step(f"episode 2 step {t}")


def test_search_task_surface_and_chain_to_orient(sscene, tmp_path):
    """BlockAssemblySearch behind RLgamesVecTaskPython; the heaps it banks are what BlockAssemblyOrient samples (SE:1348-1352 ->
    OR:419-420), in the reference's pickle layout and directly on the device"""
    import pickle
    from seqdex_b200 import bank_io
    from seqdex_b200.tasks import BlockAssemblyOrient, BlockAssemblySearch
    from seqdex_b200.vec_task import RLgamesVecTaskPython
    cfg = {"env": {"numEnvs": 32, "episodeLength": 75, "actionsMovingAverage": 0.6}, "sim": {"substeps": 2, "physx": {}}, "task": {"randomize": False}}
    task = BlockAssemblySearch(cfg, record_heaps=8)
    env = RLgamesVecTaskPython(task, "cuda:0")
    assert env.num_obs == 186 and env.num_states == 564 and env.num_actions == 23
    obs = env.reset()
    assert tuple(obs["obs"].shape) == (32, 186)
    for _ in range(75):
        o, r, d, _ = env.step(torch.rand(32, 23, device="cuda") * 2 - 1)
    assert int(task.progress_buf[0]) == 2 and torch.isfinite(r).all()      # reset() + 73 steps end the episode, step 74 resets, step 75
    tv = task.tvalue
    assert tuple(tv.shape) == (32,) and bool(((tv > 0) & (tv < 1)).all())
    task.env.tensor("SEG")[:, 0] = 500                        # make every heap count as 'target dug out' for the hand-off below
    task.reset_buf.fill_(1)
    env.step(torch.zeros(32, 23, device="cuda"))
    bank = bank_io.search_bank_valid(task.env)
    assert bank.shape[0] == 8 and bank.shape[1] >= 4 and bank.shape[2:] == (72, 13)
    bank_io.save_search_bank(task.env, task.scene, tmp_path / "heaps.pkl", tmp_path / "hands.pkl")
    with open(tmp_path / "heaps.pkl", "rb") as f:
        heaps = pickle.load(f)
    with open(tmp_path / "hands.pkl", "rb") as f:
        hands = pickle.load(f)
    assert len(heaps) == 8 and tuple(heaps[0].shape) == (11024, 132, 13) and tuple(hands[0].shape) == (11024, 23, 2)   # SE:319-332
    ocfg = {"env": {"numEnvs": 32, "episodeLength": 75, "actionsMovingAverage": 0.2}, "sim": {"substeps": 2, "physx": {}}, "task": {"randomize": False}}
    orient = BlockAssemblyOrient(ocfg, heap_bank=bank)
    orient.step(torch.zeros(32, 23, device="cuda"))
    assert int(orient.progress_buf[0]) == 1 and torch.isfinite(orient.rew_buf).all()
