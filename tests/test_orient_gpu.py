"""BlockAssemblyOrient on the GPU (csrc/sdx_task_orient.cuh through the C-ABI) against the CPU oracle: BIT-EXACT on the same
seeded inputs -- per-kernel against the golden inputs of the reference's own Python, and whole VecTask.step() trajectories
across the scripted reset (OR:1390-1695: 53 contact steps on the first reset, 103 + heap banking afterwards)."""
import os

import numpy as np
import pytest
import torch

from tests.util import lattice_bank

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def oscene():
    from seqdex_b200.tasks.cfg import scene_from_cfg
    return scene_from_cfg("BlockAssemblyOrient")   # the yaml-stated sim / env parameters (contact_offset 0.02)


def _cmp(name, a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        raise AssertionError(f"{name}: {len(bad)} of {a.size} differ; max abs diff {d.max():.3e} first at {bad[0]} "
                             f"gpu={a[tuple(bad[0])]!r} oracle={b[tuple(bad[0])]!r}")


def _pair(oscene, oracle_lib, n, wts=None):
    from seqdex_b200.env import SdxEnv
    g, o = SdxEnv(oscene, n), oracle_lib.OracleEnv(oscene, n)
    w = oracle_lib.default_tvalue_weights(1) if wts is None else wts
    g.set_tvalue_weights(w)
    o.tv = w.astype(np.float32).copy()
    return g, o


def _all(g, o, tag):
    torch.cuda.synchronize()
    for name, ov in (("OBS", o.obs), ("STATES", o.states), ("REW", o.rew), ("RESET", o.reset), ("PROGRESS", o.progress), ("TVALUE", o.tvalue),
                     ("BRICK", o.brick), ("DOF", o.dof), ("LINK", o.link), ("JAC7", o.jac7), ("TARGET_INIT", o.target_init),
                     ("SLEEP", o.slp), ("EPISODE", o.episode), ("CONSEC", o.consec), ("SUCCESSES", o.successes)):
        _cmp(f"{tag}: {name}", g.tensor(name), ov)


def test_orient_tensor_shapes(oscene):
    from seqdex_b200.env import SdxEnv
    g = SdxEnv(oscene, 8)
    assert tuple(g.tensor("OBS").shape) == (8, 186) and tuple(g.tensor("STATES").shape) == (8, 564)      # OR:206-208


def test_orient_post_and_pre_physics_kernels_on_golden_inputs(oscene, oracle_lib):
    """the inputs the reference's own Python was run on (tests/golden/orient_*.npz): GPU == oracle bit for bit, and both within
    the stated fp32 tolerance of the reference's outputs"""
    d = dict(np.load(os.path.join(G, "orient_post_physics.npz")))
    n = len(d["progress"])
    g, o = _pair(oscene, oracle_lib, n, d["tv_weights"])
    root = d["root"].reshape(n, 142, 13)
    o.set_brick_roots(np.ascontiguousarray(root[:, 9:81]))
    o.link[:] = d["rb"][:, :24]
    o.dof[:, 0, :23] = d["dof_state"][..., 0]
    o.dof[:, 1, :23] = d["dof_state"][..., 1]
    o.actions[:] = d["actions"]
    o.target_init[:, 0:3] = d["init_pos"]; o.target_init[:, 3:7] = d["init_rot"]
    o.progress[:] = d["progress"] - 1
    o.reset[:] = d["reset_in"]
    o.obs[:] = d["prev_obs"]; o.states[:] = d["prev_states"]
    o.successes[:] = d["successes"]; o.consec[:] = d["consec_in"]
    for name, v in (("BRICK", o.brick), ("LINK", o.link), ("DOF", o.dof), ("ACTIONS", o.actions), ("TARGET_INIT", o.target_init),
                    ("PROGRESS", o.progress), ("RESET", o.reset), ("OBS", o.obs), ("STATES", o.states), ("SUCCESSES", o.successes),
                    ("CONSEC", o.consec)):
        g.tensor(name).copy_(torch.from_numpy(v))
    g.post_physics()
    o.post_physics()
    torch.cuda.synchronize()
    for name, ov in (("OBS", o.obs), ("STATES", o.states), ("REW", o.rew), ("RESET", o.reset), ("PROGRESS", o.progress),
                     ("TVALUE", o.tvalue), ("CONSEC", o.consec)):
        _cmp(name, g.tensor(name), ov)
    np.testing.assert_allclose(g.tensor("OBS").cpu().numpy(), d["obs"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(g.tensor("STATES").cpu().numpy(), d["states"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(g.tensor("REW").cpu().numpy(), d["rew"], rtol=2e-5, atol=1e-7)
    assert np.array_equal(g.tensor("RESET").cpu().numpy(), d["reset"])
    # pre_physics_step on ITS golden inputs
    p = dict(np.load(os.path.join(G, "orient_pre_physics.npz")))
    g, o = _pair(oscene, oracle_lib, n)
    o.set_heap_bank(lattice_bank(oscene, 1)); g.set_heap_bank(lattice_bank(oscene, 1))
    o.dof[:, 0, :23] = p["dof_pos"]; o.dof[:, 2, :23] = p["prev_targets"]
    o.link[:, 7, 0:7] = p["hand_pose"]
    o.jac7[:] = p["jac7"]
    o.progress[:] = p["progress"]
    o.target_init[:, 0:3] = p["init_pos"]
    o.reset[:] = 0
    rows = o.brick_roots()
    for e in range(n):
        rows[e, oscene.target_brick_index(e), 0:3] = p["target_pos"][e]
    o.set_brick_roots(rows)
    for name, v in (("BRICK", o.brick), ("LINK", o.link), ("DOF", o.dof), ("JAC7", o.jac7), ("PROGRESS", o.progress),
                    ("TARGET_INIT", o.target_init), ("RESET", o.reset)):
        g.tensor(name).copy_(torch.from_numpy(v))
    a = p["actions"] * 1.3                       # some actions outside [-1, 1]: the kernel clamps like VecTask (VR:166)
    g.pre_physics(torch.from_numpy(a).cuda())
    o.pre_physics(a)
    torch.cuda.synchronize()
    assert g.last_reset_sim_steps() == 0
    _cmp("targets", g.tensor("DOF"), o.dof)
    _cmp("actions", g.tensor("ACTIONS"), o.actions)


def test_orient_steps_and_scripted_resets_bit_exact(oscene, oracle_lib):
    n = 16
    g, o = _pair(oscene, oracle_lib, n)
    bank = lattice_bank(oscene, 3)
    g.set_heap_bank(bank); o.set_heap_bank(bank)
    g.enable_orient_heap_bank(4); o.enable_orient_heap_bank(4)
    rng = np.random.default_rng(0)

    def step(tag):
        a = rng.uniform(-1, 1, size=(n, 23)).astype(np.float32)
        g.step(torch.from_numpy(a).cuda())
        o.step(a)
        assert g.last_reset_sim_steps() == o.last_reset_sim_steps, tag
        _all(g, o, tag)

    step("first reset")                      # 53 contact steps inside reset_idx, no banking (total_steps == 0)
    assert o.last_reset_sim_steps == 53
    for i in range(3):
        step(f"step {i}")
    assert o.last_reset_sim_steps == 0
    # open the gate (bias of the 'feasible' logit), time the episode out -> lockstep reset with lift + banking
    o.tv[-1] += 50.0
    g.set_tvalue_weights(o.tv)
    o.progress[:] = 73
    g.tensor("PROGRESS").fill_(73)
    step("timeout")
    assert o.reset.all()
    step("second reset")
    assert o.last_reset_sim_steps == 103 and o.ob_index.sum() > 0
    rows, index = g.orient_heap_bank()
    _cmp("banked heap rows", rows, o.ob_rows)
    _cmp("bank ring index", index, o.ob_index)
    # a PARTIAL reset: only some envs flagged; the script still steps every env (OR:1458 simulate is global)
    flags = (np.arange(n) % 3 == 0).astype(np.int64)
    o.reset[:] = flags
    g.tensor("RESET").copy_(torch.from_numpy(flags))
    step("partial reset")
    assert o.last_reset_sim_steps == 103
    assert (o.progress[flags == 1] == 1).all() and (o.progress[flags == 0] == 2).all()
    step("after partial reset")


def test_orient_bank_ring_wraps_like_the_sequential_loop(oscene, oracle_lib):
    """more banked envs than ring slots in ONE call: the parallel slot assignment must leave what the reference's sequential
    loop leaves (index += 1; if index > wrap: index = 0 -- OR:1476-1479)"""
    n = 64                                    # 8 envs per brick type, ring of wrap + 1 = 3 slots
    g, o = _pair(oscene, oracle_lib, n)
    bank = lattice_bank(oscene, 2)
    g.set_heap_bank(bank); o.set_heap_bank(bank)
    g.enable_orient_heap_bank(2); o.enable_orient_heap_bank(2)
    o.tv[-1] += 50.0
    g.set_tvalue_weights(o.tv)
    a = np.zeros((n, 23), np.float32)
    g.step(torch.from_numpy(a).cuda()); o.step(a)
    o.reset[:] = 1
    g.tensor("RESET").fill_(1)
    g.step(torch.from_numpy(a).cuda()); o.step(a)
    rows, index = g.orient_heap_bank()
    torch.cuda.synchronize()
    assert o.ob_index.tolist()[0] == 8 % 3                   # type 0: all 8 envs banked -> 8 writes into 3 slots, index wrapped twice
    assert all(np.abs(o.ob_rows[0, s]).sum() > 0 for s in range(3))
    _cmp("bank ring index", index, o.ob_index)
    _cmp("banked heap rows", rows, o.ob_rows)


def test_orient_task_and_ppo_smoke(oscene):
    """the reference-facing surface: BlockAssemblyOrient behind RLgamesVecTaskPython, one PPO iteration on 186-d observations"""
    import math
    from seqdex_b200.ppo import A2CAgent, PPOConfig
    from seqdex_b200.tasks import BlockAssemblyOrient
    from seqdex_b200.vec_task import RLgamesVecTaskPython
    cfg = {"env": {"numEnvs": 256, "episodeLength": 75, "actionsMovingAverage": 0.2}, "sim": {"substeps": 2, "physx": {}}, "task": {"randomize": False}}
    task = BlockAssemblyOrient(cfg, heap_bank=lattice_bank(oscene, 2))
    env = RLgamesVecTaskPython(task, "cuda:0")
    assert env.num_obs == 186 and env.num_states == 564 and env.num_actions == 23
    agent = A2CAgent(env, PPOConfig(minibatch_size=1024))
    info = agent.train_epoch()
    assert all(math.isfinite(v) for v in info.values()), info
    assert 0.0 < info["mean_reward"] <= 1.0                  # exp(-(5 z + 5 d)) (OR:1893)
    assert int(task.progress_buf[0]) == 9                    # reset() step + 8 rollout steps


def test_chain_hand_off_orient_to_grasp_sim(oscene, scene, tmp_path):
    """the link of the chain GraspSim depends on (GS:412-413): the heaps Orient leaves face up, in the reference's pickle layout
    (list[8] of Tensor[11024, 132, 13]) and directly on the device, become the bank GraspSim samples on reset"""
    import pickle
    from seqdex_b200 import bank_io
    from seqdex_b200.env import SdxEnv
    from seqdex_b200.tasks.block_assembly_grasp_sim import default_tvalue_weights
    n = 32
    w = default_tvalue_weights(1)
    w[-1] += 50.0                                            # gate open: every heap with the brick in the near half is banked
    g = SdxEnv(oscene, n)
    g.set_tvalue_weights(w)
    g.set_heap_bank(lattice_bank(oscene, 2))
    g.enable_orient_heap_bank(16)
    a = torch.zeros(n, 23, device="cuda")
    g.step(a)
    for _ in range(2):                                       # two lockstep resets -> up to 8 heaps per brick type
        g.tensor("RESET").fill_(1)
        g.step(a)
    bank = bank_io.orient_bank_valid(g)
    assert bank.shape[0] == 8 and bank.shape[2:] == (72, 13) and bank.shape[1] >= 2 and float(bank[..., 7:13].abs().max()) == 0
    p = tmp_path / "saved_searching_ternimal_states_good_mo_tvalue.pkl"
    bank_io.save_orient_heap_bank(g, oscene, p)
    with open(p, "rb") as f:
        ref = pickle.load(f)
    assert len(ref) == 8 and all(tuple(t.shape) == (11024, 132, 13) for t in ref)
    k = bank.shape[1]
    for ty in range(8):
        assert torch.equal(ref[ty][:k, :72, 0:7], bank[ty].cpu()[..., 0:7])
        assert torch.equal(ref[ty][0, 72:], torch.from_numpy(np.ctypeslib.as_array(oscene.c.fixed_root).reshape(60, 13).copy()))
    back = bank_io.load_heap_bank(p)                         # what GraspSim's loader makes of the file: K = 11024 rows, zeros beyond
    assert torch.equal(back[:, :k], bank.cpu())
    # next stage: GraspSim resets from exactly these heaps
    gs = SdxEnv(scene, n)
    gs.set_tvalue_weights(default_tvalue_weights(1))
    gs.set_heap_bank(bank)
    gs.pre_physics(a)                                        # reset_idx (all reset flags start at 1, BT:63)
    roots = gs.brick_roots().cpu()
    for e in range(n):
        match = [torch.equal(roots[e, :, 0:7], bank[e % 8, s].cpu()[:, 0:7]) for s in range(k)]
        # root row -> COM block -> root row round-trips to the last bit only for axis-aligned bricks; allow 1e-6
        close = [float((roots[e, :, 0:7] - bank[e % 8, s].cpu()[:, 0:7]).abs().max()) < 1e-6 for s in range(k)]
        assert any(match) or any(close), e
