"""ToolPositioningGrasp / ToolPositioningOrient on the GPU (csrc/sdx_task_tool.cuh through the C-ABI) against the CPU oracle, which
tests/test_tool_oracle_golden.py pins to the reference's own Python: bit-exact on the golden inputs and over whole episodes (reset,
banking, contact step with the two-box hammer, observations with history, reward, resets)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NA = 11
TASKS = {"grasp": "ToolPositioningGrasp", "orient": "ToolPositioningOrient"}


def _scene(name):
    from seqdex_b200.tasks.cfg import scene_from_cfg
    return scene_from_cfg(TASKS[name])


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
                     ("PLATE", o.plate), ("SUCCESS", o.success_buf), ("NCONTACT", o.ncontact), ("SUCCESSES", o.successes), ("CONSEC", o.consec)):
        _cmp(f"{tag}: {name}", g.tensor(name), ov)
    h, ob, ix = g.grasp_bank()
    _cmp(f"{tag}: grasp ring index", ix, o.gb_index)
    _cmp(f"{tag}: grasp ring hand", h, o.gb_hand)
    _cmp(f"{tag}: grasp ring obj", ob, o.gb_obj)


def _rows72(tool_rows):
    out = np.zeros((tool_rows.shape[0], 72, 13), np.float32)
    out[..., 6] = 1
    out[:, 0] = tool_rows
    return out


def test_tool_tensor_shapes_and_missing_bank():
    from seqdex_b200.env import SdxEnv
    g = SdxEnv(_scene("grasp"), 8)
    assert tuple(g.tensor("OBS").shape) == (8, 468) and tuple(g.tensor("STATES").shape) == (8, 564)      # TG:238-243
    g.step(torch.zeros(8, 23, device="cuda"))           # Grasp needs no bank
    g2 = SdxEnv(_scene("orient"), 8)
    with pytest.raises(RuntimeError, match="banked grasps"):
        g2.step(torch.zeros(8, 23, device="cuda"))


@pytest.mark.parametrize("name", ["grasp", "orient"])
def test_tool_post_physics_kernel_on_golden_inputs(name, oracle_lib):
    """the inputs the reference's own Python was run on, two calls in a row: GPU == oracle bit for bit, and within the stated fp32
    tolerance of the reference's outputs"""
    from seqdex_b200.env import SdxEnv
    d = dict(np.load(os.path.join(G, f"tool_{name}_post.npz")))
    n = len(d["progress0"])
    sc = _scene(name)
    g, o = SdxEnv(sc, n), oracle_lib.OracleEnv(sc, n)
    o.target_init[:, 0:3] = d["init_pos"]; o.target_init[:, 3:7] = d["init_rot"]
    o.obs[:, 0:312] = d["hist_obs"].reshape(n, 312)
    o.states[:, 0:376] = d["hist_states"].reshape(n, 376)
    o.consec[:] = d["consec_in"]
    o.reset[:] = d["reset_in0"]
    o.successes[:] = d["successes_in0"]
    for name_, src in (("TARGET_INIT", o.target_init), ("OBS", o.obs), ("STATES", o.states), ("CONSEC", o.consec), ("RESET", o.reset),
                       ("SUCCESSES", o.successes)):
        g.tensor(name_).copy_(torch.from_numpy(np.ascontiguousarray(src)))
    for call in (0, 1):
        root = d[f"root{call}"].reshape(n, NA, 13)
        o.set_brick_roots(_rows72(root[:, 9]))
        o.link[:] = d[f"rb{call}"][:, :24]
        o.dof[:, 0, :23] = d[f"dof_state{call}"][..., 0]
        o.dof[:, 1, :23] = d[f"dof_state{call}"][..., 1]
        o.actions[:] = d[f"actions{call}"]
        o.plate[:] = root[:, 10, 0:7]
        o.progress[:] = d[f"progress{call}"] - 1
        for name_, src in (("BRICK", o.brick), ("LINK", o.link), ("DOF", o.dof), ("ACTIONS", o.actions), ("PLATE", o.plate), ("PROGRESS", o.progress)):
            g.tensor(name_).copy_(torch.from_numpy(np.ascontiguousarray(src)))
        g.post_physics(); o.post_physics()
        torch.cuda.synchronize()
        for name_, ov in (("OBS", o.obs), ("STATES", o.states), ("REW", o.rew), ("RESET", o.reset), ("PROGRESS", o.progress), ("SUCCESSES", o.successes),
                          ("CONSEC", o.consec)):
            _cmp(f"call {call}: {name_}", g.tensor(name_), ov)
        np.testing.assert_allclose(g.tensor("OBS").cpu().numpy(), d[f"obs{call}"], rtol=0, atol=3e-6)
        np.testing.assert_allclose(g.tensor("STATES").cpu().numpy(), d[f"states{call}"], rtol=0, atol=3e-6)
        np.testing.assert_allclose(g.tensor("REW").cpu().numpy(), d[f"rew{call}"], rtol=3e-5, atol=2e-7)
        assert np.array_equal(g.tensor("RESET").cpu().numpy(), d[f"reset{call}"])
        np.testing.assert_array_equal(g.tensor("SUCCESSES").cpu().numpy(), d[f"successes{call}"])
        np.testing.assert_allclose(g.tensor("CONSEC").cpu().numpy(), d[f"consec{call}"], rtol=1e-5)


def test_tool_grasp_reset_kernel_on_golden_inputs(oracle_lib):
    """reset_idx executed by the reference (tests/golden/tool_grasp_reset.npz) vs k_tool_bank + k_tool_reset with the same pitch / yaw draws"""
    from seqdex_b200.env import SdxEnv
    d = dict(np.load(os.path.join(G, "tool_grasp_reset.npz")))
    root = d["root"].reshape(-1, NA, 13)
    n = root.shape[0]
    sc = _scene("grasp")
    g, o = SdxEnv(sc, n), oracle_lib.OracleEnv(sc, n)
    g.step(torch.zeros(n, 23, device="cuda"))          # one ordinary step first: banking / success_buf only once total_steps > 0 (TG:1436)
    o.set_brick_roots(_rows72(root[:, 9]))
    g.tensor("BRICK").copy_(torch.from_numpy(o.brick))
    dof = np.zeros((n, 3, 24), np.float32)
    dof[:, 0, :23] = d["dof_state"][..., 0]; dof[:, 1, :23] = d["dof_state"][..., 1]
    g.tensor("DOF").copy_(torch.from_numpy(dof))
    g.tensor("PLATE").copy_(torch.from_numpy(np.ascontiguousarray(root[:, 10, 0:7])))
    g.tensor("PROGRESS").copy_(torch.from_numpy(d["progress"]))
    g.tensor("SUCCESSES").copy_(torch.from_numpy(d["successes"]))
    g.aux()[1].copy_(torch.from_numpy(d["finger_dist"]))
    g.tensor("OBS").fill_(1.5); g.tensor("STATES").fill_(-2.5)
    rs = np.zeros(n, np.int64); rs[d["env_ids"]] = 1
    g.tensor("RESET").copy_(torch.from_numpy(rs))
    h, ob, ix = g.grasp_bank()
    h.zero_(); ob.zero_()
    ix.copy_(torch.from_numpy(d["index_in"].astype(np.int32)))
    g.tool_test_hooks(None, int(d["pitch_k"]), d["yaw_u"])
    g.pre_physics(torch.zeros(n, 23, device="cuda"))
    torch.cuda.synchronize()
    ids = d["env_ids"]
    rest = np.setdiff1d(np.arange(n), ids)
    out = d["root_out"].reshape(n, NA, 13)
    got = g.brick_roots().cpu().numpy()[:, 0]
    np.testing.assert_allclose(got[ids], out[ids, 9], rtol=0, atol=2e-6)
    np.testing.assert_allclose(g.tensor("PLATE").cpu().numpy()[ids], out[ids, 10, 0:7], rtol=0, atol=1e-7)
    gd = g.tensor("DOF").cpu().numpy()
    np.testing.assert_allclose(gd[ids, 0, :23], d["dof_out"][ids, :, 0], rtol=0, atol=1e-7)
    assert float(np.abs(gd[ids, 1, :23]).max()) == 0.0
    np.testing.assert_allclose(g.tensor("TARGET_INIT").cpu().numpy()[ids, 0:3], d["init_pos"][ids], atol=1e-7)
    np.testing.assert_array_equal(g.tensor("SUCCESS").cpu().numpy()[ids, 0], d["success_buf"][ids, 0])
    assert np.array_equal(g.tensor("PROGRESS").cpu().numpy(), d["progress_out"]) and np.array_equal(g.tensor("RESET").cpu().numpy(), d["reset_out"])
    obs, st = g.tensor("OBS").cpu().numpy(), g.tensor("STATES").cpu().numpy()
    assert float(np.abs(obs[ids]).max()) == 0.0 and float(np.abs(st[ids]).max()) == 0.0 and np.all(obs[rest] == 1.5) and np.all(st[rest] == -2.5)
    np.testing.assert_array_equal(ix.cpu().numpy(), d["index_out"])
    where = d["bank_where"]
    np.testing.assert_array_equal(h.cpu().numpy()[where[:, 0], where[:, 1]], d["bank_hand_rows"])
    np.testing.assert_allclose(ob.cpu().numpy()[where[:, 0], where[:, 1]], d["bank_obj_rows"], rtol=0, atol=2e-6)
    mask = np.zeros((8, 11024), bool); mask[where[:, 0], where[:, 1]] = True
    assert float(np.abs(ob.cpu().numpy()[~mask]).max()) == 0.0


def test_tool_orient_reset_kernel_on_golden_inputs(oracle_lib):
    from seqdex_b200.env import SdxEnv
    d = dict(np.load(os.path.join(G, "tool_orient_reset.npz")))
    root = d["root"].reshape(-1, NA, 13)
    n = root.shape[0]
    sc = _scene("orient")
    g, o = SdxEnv(sc, n), oracle_lib.OracleEnv(sc, n)
    g.set_grasp_bank(d["bank_hand"], d["bank_obj"])
    g.step(torch.zeros(n, 23, device="cuda"))
    o.set_brick_roots(_rows72(root[:, 9]))
    g.tensor("BRICK").copy_(torch.from_numpy(o.brick))
    dof = np.zeros((n, 3, 24), np.float32)
    dof[:, 0, :23] = d["dof_state"][..., 0]; dof[:, 1, :23] = d["dof_state"][..., 1]
    g.tensor("DOF").copy_(torch.from_numpy(dof))
    g.tensor("PLATE").copy_(torch.from_numpy(np.ascontiguousarray(root[:, 10, 0:7])))
    g.tensor("PROGRESS").copy_(torch.from_numpy(d["progress"]))
    rs = np.zeros(n, np.int64); rs[d["env_ids"]] = 1
    g.tensor("RESET").copy_(torch.from_numpy(rs))
    g.tool_test_hooks(d["slot_by_env"], -1, None)
    g.pre_physics(torch.zeros(n, 23, device="cuda"))
    torch.cuda.synchronize()
    ids = d["env_ids"]
    out = d["root_out"].reshape(n, NA, 13)
    got = g.brick_roots().cpu().numpy()[:, 0]
    np.testing.assert_allclose(got[ids], out[ids, 9], rtol=0, atol=3e-6)
    gd = g.tensor("DOF").cpu().numpy()
    np.testing.assert_array_equal(gd[ids, 0, :23], d["dof_out"][ids, :, 0])
    np.testing.assert_array_equal(gd[ids, 1, :23], d["dof_out"][ids, :, 1])
    np.testing.assert_array_equal(g.tensor("TARGET_INIT").cpu().numpy()[ids, 0:3], d["init_pos"][ids])
    np.testing.assert_array_equal(g.tensor("SUCCESS").cpu().numpy()[ids, 0], d["success_buf"][ids, 0])
    assert np.array_equal(g.tensor("PROGRESS").cpu().numpy(), d["progress_out"]) and np.array_equal(g.tensor("RESET").cpu().numpy(), d["reset_out"])


@pytest.mark.parametrize("name,n,steps", [("grasp", 6, 170), ("grasp", 24, 170), ("orient", 6, 140), ("orient", 24, 140)])
def test_tool_whole_episodes_bit_exact(name, n, steps, oracle_lib):
    """reset_idx (Grasp: drawn pitch / yaw, banking; Orient: banked grasps), pre-physics, contact step with the compound hammer,
    observations with history, reward, resets: more than one episode, so every env passes through a time-out reset"""
    from seqdex_b200.env import SdxEnv
    from seqdex_b200.tasks.tool_positioning import synthetic_tool_grasp_bank
    sc = _scene(name)
    g, o = SdxEnv(sc, n), oracle_lib.OracleEnv(sc, n)
    if name == "orient":
        hand, obj = synthetic_tool_grasp_bank(sc, 3, seed=5)
        g.set_grasp_bank(hand, obj); o.set_grasp_bank(hand, obj)
    rng = np.random.default_rng(11)
    resets, contacts, pitches = 0, 0, set()
    for t in range(steps):
        a = rng.uniform(-1, 1, size=(n, 23)).astype(np.float32) * (0.3 if t % 50 < 25 else 1.0)
        g.step(torch.from_numpy(a).cuda()); o.step(a)
        resets += int(o.reset.sum())
        contacts += int(o.ncontact[:, 0].sum())
        if t % 10 == 9 or t < 3:
            _all(g, o, f"step {t}")
    _all(g, o, "end")
    assert resets >= n and contacts > 0 and np.isfinite(o.brick[:, :, :1]).all()


def test_tool_grasp_scripted_lift_banks_grasps():
    """a hand-written open-loop policy -- reach down over the handle, close, let the script lift (TG:1622-1626) -- has to get SOME
    hammers off the bin floor with the two-box model; whatever passes the gate lands in the rings"""
    from seqdex_b200.tasks import ToolPositioningGrasp
    t = ToolPositioningGrasp({"env": {"numEnvs": 64}, "sim": {}, "task": {"randomize": False}})
    a = torch.zeros(64, 23, device="cuda")
    z0 = None
    zmax = torch.zeros(64, device="cuda")
    for k in range(149):
        a.zero_()
        if k < 25:
            a[:, 2] = -1.0; a[:, 7:] = -1.0          # descend with the hand open
        else:
            a[:, 7:] = 1.0                           # close and hold
        t.step(a)
        z = t.env.brick_roots()[:, 0, 2]
        z0 = z.clone() if z0 is None else z0
        zmax = torch.maximum(zmax, z)
    assert torch.isfinite(t.obs_buf).all() and torch.isfinite(t.rew_buf).all()
    assert float(t.rew_buf.min()) >= 0.0 and float(t.rew_buf.max()) <= 4.0 + 1e-5        # exp(...) <= 1, x (1 + 10 * 0.2) + 1 (TG:1877)
    assert float(z.min()) > 0.55, "no hammer may fall through the bin / table"


def test_tool_orient_online_tvalue_update_vs_reference(oracle_lib):
    """the online t-value update of TO's reset_idx (TO:1305-1350) executed by the reference with `if_t_value` on: same labels (kernel ==
    oracle == reference), and five Adam steps on the tensor-core MLP follow the reference's five losses within bf16 tolerance"""
    from seqdex_b200.tasks import ToolPositioningOrient
    d = dict(np.load(os.path.join(G, "tool_orient_reset.npz")))
    n = d["tv_obs_in"].shape[0]
    t = ToolPositioningOrient({"env": {"numEnvs": n}, "sim": {}, "task": {"randomize": False}}, if_t_value=True)
    o = oracle_lib.OracleEnv(t.scene, n)
    rows = np.zeros((n, 13), np.float32)
    rows[:, 0:3], rows[:, 3:7] = d["tv_target_pos"], d["tv_target_rot"]
    o.set_brick_roots(_rows72(rows))
    o.plate[:] = d["tv_plate"]
    t.env.tensor("BRICK").copy_(torch.from_numpy(o.brick))
    t.env.tensor("PLATE").copy_(torch.from_numpy(o.plate))
    t.segmentation_target_init.copy_(torch.from_numpy(d["tv_obs_in"]))
    t.t_value.load_flat(torch.from_numpy(d["tv_w0"]))
    losses = []
    for k in range(5):
        losses.append(float(t.online_t_value_update(steps=1)))
    torch.cuda.synchronize()
    _cmp("labels", t._tv_label, o.tool_tvalue_labels())
    _cmp("success_buf", t.success_buf, o.success_buf)
    np.testing.assert_array_equal(t.success_buf.cpu().numpy(), d["tv_success_buf"])
    np.testing.assert_allclose(losses, d["tv_losses"], rtol=0, atol=3e-3)
    assert losses[-1] < losses[0]
    dw, dref = t.t_value.params.cpu() - torch.from_numpy(d["tv_w0"]), torch.from_numpy(d["tv_w5"] - d["tv_w0"])
    big = dref.abs() > 1.2e-3                                                # five Adam steps at lr 3e-4 with a steady gradient sign move 1.5e-3
    assert int(big.sum()) > 1000
    assert float((torch.sign(dw[big]) == torch.sign(dref[big])).float().mean()) > 0.97
    assert float((dw - dref).abs().mean()) < 2e-4


def test_tool_orient_runs_with_online_tvalue():
    """with `if_t_value` the update runs inside step() whenever reset_idx is about to (TO:1305), and success_buf then holds its labels"""
    from seqdex_b200.tasks import ToolPositioningOrient
    t = ToolPositioningOrient({"env": {"numEnvs": 64}, "sim": {}, "task": {"randomize": False}}, if_t_value=True)
    w0 = t.t_value.params.clone()
    g = torch.Generator(device="cuda").manual_seed(3)
    for k in range(130):                                                      # one 125-step episode: every env times out once
        t.step(torch.rand(64, 23, device="cuda", generator=g) * 2 - 1)
    assert "BCE_loss" in t.extras and torch.isfinite(t.extras["BCE_loss"])
    assert float((t.t_value.params - w0).abs().max()) > 0
    sb = t.success_buf
    assert torch.equal(sb[:, 1], (sb[:, 0] <= 0.5).float())                   # TO:1316


def test_tool_chain_insertion_obs_and_inner_step(oracle_lib):
    """ToolPositioningChain's two C-ABI pieces against the oracle: compute_insertion_observations on the golden inputs (== the reference's
    output exactly), and the inner loop's step (fingers from the actions, arm holds, contact step) bit for bit over 20 steps"""
    from seqdex_b200.env import SdxEnv
    d = dict(np.load(os.path.join(G, "tool_grasp_post.npz")))
    n = len(d["progress0"])
    sc = _scene("grasp")
    g, o = SdxEnv(sc, n), oracle_lib.OracleEnv(sc, n)
    g.tensor("OBS").copy_(torch.from_numpy(d["obs1"]))
    ins = torch.zeros(n, 468, device="cuda")
    ins[:, 0:312] = torch.from_numpy(d["ins_hist"].reshape(n, 312)).cuda()
    g.tool_insertion_obs(torch.from_numpy(d["ins_actions"]).cuda(), torch.from_numpy(d["ins_progress"]).cuda(), ins)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ins.cpu().numpy(), d["ins_obs"])
    rng = np.random.default_rng(2)
    a = rng.uniform(-1, 1, size=(n, 23)).astype(np.float32)
    g.step(torch.from_numpy(a).cuda()); o.step(a)                        # every env reset once, the hammer falls
    acts_before = g.tensor("ACTIONS").clone()
    for k in range(20):
        a = rng.uniform(-1, 1, size=(n, 23)).astype(np.float32)
        g.tool_inner_step(torch.from_numpy(a).cuda()); o.tool_inner_step(a)
    torch.cuda.synchronize()
    for name, ov in (("BRICK", o.brick), ("DOF", o.dof), ("LINK", o.link), ("NCONTACT", o.ncontact), ("PROGRESS", o.progress)):
        _cmp(f"inner: {name}", g.tensor(name), ov)
    assert torch.equal(g.tensor("ACTIONS"), acts_before)                  # the task's own actions tensor is untouched by the inner loop


def test_tool_chain_task_runs_the_inner_policy_at_step_118():
    from seqdex_b200.tasks import ToolPositioningChain
    n = 32
    t = ToolPositioningChain({"env": {"numEnvs": n}, "sim": {}, "task": {"randomize": False}})
    gen = torch.Generator(device="cuda").manual_seed(5)
    launches = []
    for k in range(125):
        l0 = t.env.launch_count()
        t.step((torch.rand(n, 23, device="cuda", generator=gen) * 2 - 1) * 0.2)
        launches.append(t.env.launch_count() - l0)
    # episodes run in lockstep under small actions (no early reset), so env 0's clock reads 118 at the 119th call
    assert t.inner_calls == 1
    assert launches[118] >= 2 * 125 and max(launches[:118]) < 20           # 125 x (targets + contact step) inside ONE outer step (TC:1734)
    assert int(t.insertion_progress_buf.max()) == 1                       # zeroed inside the loop, incremented at its end (TC:1739, 1768)
    assert torch.isfinite(t.insertion_obs_buf).all() and torch.isfinite(t.obs_buf).all()
    np.testing.assert_array_equal(t.insertion_obs_buf[:, 23:46].cpu().numpy(), t.insertion_actions.cpu().numpy())
    assert torch.equal(t.insertion_obs_buf[:, 0:23], t.obs_buf[:, 0:23]) and torch.equal(t.insertion_obs_buf[:, 61:156], t.obs_buf[:, 61:156])
