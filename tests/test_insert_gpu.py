"""BlockAssemblyInsertSim on the GPU (csrc/sdx_task_insert.cuh through the C-ABI) against the CPU oracle, which
tests/test_insert_oracle_golden.py pins to the reference's own Python: bit-exact on the golden inputs and over whole episodes
(reset from the banked grasps, contact step with the env-selective base-plate, observations, reward, resets) at 6 and 24 envs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NA = 18


@pytest.fixture(scope="module")
def iscene():
    from seqdex_b200.tasks.cfg import scene_from_cfg
    return scene_from_cfg("BlockAssemblyInsertSim")


def _cmp(name, a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        raise AssertionError(f"{name}: {len(bad)} of {a.size} differ; max abs diff {d.max():.3e} first at {bad[0]} "
                             f"gpu={a[tuple(bad[0])]!r} oracle={b[tuple(bad[0])]!r}")


def _all(g, o, tag):
    torch.cuda.synchronize()
    for name, ov in (("OBS", o.obs), ("STATES", o.states), ("REW", o.rew), ("RESET", o.reset), ("PROGRESS", o.progress), ("BRICK", o.brick),
                     ("DOF", o.dof), ("LINK", o.link), ("JAC7", o.jac7), ("TARGET_INIT", o.target_init), ("SLEEP", o.slp), ("EPISODE", o.episode),
                     ("PLATE", o.plate), ("ROT_ERR", o.rot_err), ("SUCCESS", o.success_buf), ("NCONTACT", o.ncontact), ("SUCCESSES", o.successes)):
        _cmp(f"{tag}: {name}", g.tensor(name), ov)


def _bricks72(rows8):
    out = np.zeros((rows8.shape[0], 72, 13), np.float32)
    out[..., 6] = 1
    out[:, :8] = rows8
    return out


def test_insert_tensor_shapes_and_missing_bank(iscene):
    from seqdex_b200.env import SdxEnv
    g = SdxEnv(iscene, 8)
    assert tuple(g.tensor("OBS").shape) == (8, 75) and tuple(g.tensor("STATES").shape) == (8, 188)      # IS:187-191
    with pytest.raises(RuntimeError, match="banked grasps"):
        g.step(torch.zeros(8, 23, device="cuda"))


def test_insert_post_physics_kernel_on_golden_inputs(iscene, oracle_lib):
    """the inputs the reference's own Python was run on (tests/golden/insert_post_physics.npz): GPU == oracle bit for bit, and within
    the stated fp32 tolerance of the reference's outputs"""
    from seqdex_b200.env import SdxEnv
    d = dict(np.load(os.path.join(G, "insert_post_physics.npz")))
    n = len(d["progress"])
    g, o = SdxEnv(iscene, n), oracle_lib.OracleEnv(iscene, n)
    root = d["root"].reshape(n, NA, 13)
    o.set_brick_roots(_bricks72(root[:, 9:17]))
    o.link[:] = d["rb"][:, :24]
    o.dof[:, 0, :23] = d["dof_state"][..., 0]
    o.dof[:, 1, :23] = d["dof_state"][..., 1]
    o.actions[:] = d["actions"]
    o.target_init[:, 0:3] = d["init_pos"]; o.target_init[:, 3:7] = d["init_rot"]
    o.plate[:] = root[:, 17, 0:7]
    o.rot_err[:] = d["rot_err"]
    o.progress[:] = d["progress"] - 1
    o.reset[:] = d["reset_in"]
    o.obs[:] = d["prev_obs"]; o.states[:] = d["prev_states"]
    o.successes[:] = d["successes"]; o.consec[:] = d["consec_in"]
    for name, src in (("BRICK", o.brick), ("LINK", o.link), ("DOF", o.dof), ("ACTIONS", o.actions), ("TARGET_INIT", o.target_init), ("PLATE", o.plate),
                      ("ROT_ERR", o.rot_err), ("PROGRESS", o.progress), ("RESET", o.reset), ("OBS", o.obs), ("STATES", o.states),
                      ("SUCCESSES", o.successes), ("CONSEC", o.consec)):
        g.tensor(name).copy_(torch.from_numpy(np.ascontiguousarray(src)))
    g.post_physics(); o.post_physics()
    torch.cuda.synchronize()
    for name, ov in (("OBS", o.obs), ("STATES", o.states), ("REW", o.rew), ("RESET", o.reset), ("PROGRESS", o.progress)):
        _cmp(name, g.tensor(name), ov)
    np.testing.assert_allclose(g.tensor("OBS").cpu().numpy(), d["obs"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(g.tensor("STATES").cpu().numpy(), d["states"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(g.tensor("REW").cpu().numpy(), d["rew"], rtol=3e-5, atol=2e-7)
    assert np.array_equal(g.tensor("RESET").cpu().numpy(), d["reset"])
    np.testing.assert_allclose(g.tensor("CONSEC").cpu().numpy(), d["consec"], rtol=1e-5)


def test_insert_reset_kernel_on_golden_inputs(iscene):
    """reset_idx executed by the reference (tests/golden/insert_reset.npz) vs k_insert_reset with the same slots and plate yaw"""
    from seqdex_b200.env import SdxEnv
    from seqdex_b200.scene import Scene
    d = dict(np.load(os.path.join(G, "insert_reset.npz")))
    root = d["root"].reshape(-1, NA, 13)
    n = root.shape[0]
    g = SdxEnv(iscene, n)
    g.set_grasp_bank(d["bank_hand"], d["bank_obj"])
    g.step(torch.zeros(n, 23, device="cuda"))          # one ordinary step first: success_buf is only written once total_steps > 0 (IS:1341)
    rows = _bricks72(root[:, 9:17])
    for e in range(n):
        rows[e, Scene.target_brick_index(e), 0:3] = d["seg_pos"][e]
        rows[e, Scene.target_brick_index(e), 3:7] = d["seg_rot"][e]
    from oracle import oracle as O
    o = O.OracleEnv(iscene, n)
    o.set_brick_roots(rows)
    g.tensor("BRICK").copy_(torch.from_numpy(o.brick))
    dof = np.zeros((n, 3, 24), np.float32)
    dof[:, 0, :23] = d["dof_state"][..., 0]; dof[:, 1, :23] = d["dof_state"][..., 1]
    g.tensor("DOF").copy_(torch.from_numpy(dof))
    g.tensor("PLATE").copy_(torch.from_numpy(np.ascontiguousarray(root[:, 17, 0:7])))
    g.tensor("PROGRESS").copy_(torch.from_numpy(d["progress"]))
    g.tensor("SUCCESSES").copy_(torch.from_numpy(d["successes"]))
    rs = np.zeros(n, np.int64); rs[d["env_ids"]] = 1
    g.tensor("RESET").copy_(torch.from_numpy(rs))
    slots = np.zeros(n, np.int32); slots[d["env_ids"]] = d["slots"]
    g.insert_test_hooks(slots, int(d["plate_rot"]))
    g.pre_physics(torch.zeros(n, 23, device="cuda"))
    torch.cuda.synchronize()
    ids = d["env_ids"]
    out = d["root_out"].reshape(n, NA, 13)
    got = g.brick_roots().cpu().numpy()[:, :8]
    np.testing.assert_allclose(got[ids], out[ids, 9:17], rtol=0, atol=2e-6)
    np.testing.assert_allclose(g.tensor("PLATE").cpu().numpy()[ids], out[ids, 17, 0:7], rtol=0, atol=1e-7)
    gd = g.tensor("DOF").cpu().numpy()
    np.testing.assert_array_equal(gd[ids, 0, :23], d["dof_out"][ids, :, 0])
    assert float(np.abs(gd[ids, 1, :23]).max()) == 0.0
    np.testing.assert_array_equal(g.tensor("TARGET_INIT").cpu().numpy()[ids, 0:3], d["init_pos"][ids])
    np.testing.assert_array_equal(g.tensor("SUCCESS").cpu().numpy()[ids], d["success_buf"][ids])
    assert np.array_equal(g.tensor("PROGRESS").cpu().numpy(), d["progress_out"]) and np.array_equal(g.tensor("RESET").cpu().numpy(), d["reset_out"])


@pytest.mark.parametrize("n", [6, 24])
def test_insert_whole_episodes_bit_exact(iscene, oracle_lib, n):
    """reset_idx from the grasp bank, pre-physics, contact step (base-plate by env % 3), observations / reward / resets: 140 steps
    = more than one 125-step episode, so every env passes through a time-out reset, most through an early one"""
    from seqdex_b200.env import SdxEnv
    from seqdex_b200.tasks.block_assembly_insert_sim import synthetic_grasp_bank
    g, o = SdxEnv(iscene, n), oracle_lib.OracleEnv(iscene, n)
    hand, obj = synthetic_grasp_bank(iscene, 3, seed=5)
    g.set_grasp_bank(hand, obj); o.set_grasp_bank(hand, obj)
    rng = np.random.default_rng(7)
    resets, yaws = 0, set()
    for t in range(140):
        a = rng.uniform(-1, 1, size=(n, 23)).astype(np.float32) * (0.3 if t % 50 < 25 else 1.0)
        g.step(torch.from_numpy(a).cuda()); o.step(a)
        resets += int(o.reset.sum())
        yaws |= set(np.unique(o.plate[:, 5]).tolist())
        if t % 10 == 9 or t < 3:
            _all(g, o, f"step {t}")
    _all(g, o, "end")
    assert resets >= n and np.isfinite(o.brick[:, :, :8]).all()
    assert len(yaws) == 2, "both base-plate yaws must occur over the run"
