"""Isaac-Gym-shaped tensor API (boundary B2): refresh_* materialise the reference's tensor layouts from the kernel's
internal state, set_*_indexed write caller rows back.  Checked against the oracle's root-row conversion and the scene
constants; indices are sim-domain actor indices (int32) as in GS:1501-1516, 1538-1545."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_refresh_root_rb_dof_jacobian(scene, oracle_lib):
    from seqdex_b200.env import SdxEnv
    n = 5
    g, o = SdxEnv(scene, n), oracle_lib.OracleEnv(scene, n)
    for _ in range(3):
        g.simulate(); o.simulate()
    for name in ("ROOT", "RB", "DOF_STATE", "JACOBIAN"):
        g.refresh(name)
    torch.cuda.synchronize()
    root = g.tensor("ROOT").view(n, 142, 13).cpu().numpy()
    ref = o.brick_roots()
    assert np.array_equal(g.tensor('BRICK').cpu().numpy(), o.brick), 'internal state differs'
    assert np.array_equal(root[:, 9:81], ref), f'max diff {np.abs(root[:, 9:81] - ref).max()} at {np.argwhere(root[:, 9:81] != ref)[:4]}'   # 72 free bricks, actor slots 9..80
    fixed = np.ctypeslib.as_array(scene.c.fixed_root).reshape(60, 13)
    assert np.allclose(root[:, 81:141], fixed[None])                            # 60 fixed bricks
    assert np.allclose(root[:, 0, :7], [-0.35, 0, 0.6, 0, 0, 0, 1])             # hand actor root (GS:625)
    assert np.allclose(root[:, 3, :3], [0, 0, 0.3]) and np.allclose(root[:, 141, :3], [0.25, -0.19, 0.618])
    rb = g.tensor("RB").view(n, 165, 13).cpu().numpy()
    assert np.array_equal(rb[:, :24], o.link)                                   # robot links first (rigid_body_states[:, 7] = link7)
    assert np.array_equal(rb[:, 24 + 8:24 + 8 + 72], o.brick_roots())           # actors 9.. -> bodies 32..
    ds = g.tensor("DOF_STATE").view(n, 23, 2).cpu().numpy()
    assert np.array_equal(ds[..., 0], o.dof[:, 0, :23]) and np.array_equal(ds[..., 1], o.dof[:, 1, :23])
    J = g.tensor("JACOBIAN").cpu().numpy()                                      # [N, 23, 6, 23]; task reads J[:, 7-1, :, :7] (GS:1601)
    np.testing.assert_allclose(J[:, 6, :, :7], o.jac7, rtol=0, atol=2e-6)
    assert np.all(J[:, 6, :, 7:] == 0)                                          # finger DoFs do not move link7
    assert np.all(J[:, 0, :, 1:] == 0)                                          # link1 moves with DoF 0 only


def test_set_indexed_roundtrip(scene, oracle_lib):
    from seqdex_b200 import _lib
    from seqdex_b200.env import SdxEnv
    n = 6
    g, o = SdxEnv(scene, n), oracle_lib.OracleEnv(scene, n)
    g.refresh("ROOT"); g.refresh("DOF_STATE")
    root = g.tensor("ROOT")
    rows = o.brick_roots()
    rng = np.random.default_rng(0)
    envs = [1, 4]
    for e in envs:
        rows[e, :, 0:3] += rng.uniform(-0.02, 0.02, size=(72, 3)).astype(np.float32)
        rows[e, :, 7:13] = rng.normal(size=(72, 6)).astype(np.float32) * 0.1
    root.view(n, 142, 13)[:, 9:81] = torch.from_numpy(rows).cuda()
    idx = torch.tensor(sorted({e * 142 + a for e in envs for a in range(9, 81)} | {envs[0] * 142 + 1}), dtype=torch.int32, device="cuda")
    _lib.check(g.L.sdx_set_actor_root_state_indexed(g.h, ctypes.c_void_p(root.data_ptr()), ctypes.c_void_p(idx.data_ptr()), idx.numel()))
    o.set_brick_roots(np.where(np.isin(np.arange(n), envs)[:, None, None], rows, o.brick_roots()))
    ds = g.tensor("DOF_STATE").view(n, 23, 2)
    tg = torch.zeros(n, 23, device="cuda")
    ds[2, :, 0] = 0.3; ds[2, :, 1] = -0.2; tg[2] = 0.25
    hand = torch.tensor([2 * 142], dtype=torch.int32, device="cuda")
    _lib.check(g.L.sdx_set_dof_state_indexed(g.h, ctypes.c_void_p(ds.data_ptr()), ctypes.c_void_p(hand.data_ptr()), 1))
    _lib.check(g.L.sdx_set_dof_target_indexed(g.h, ctypes.c_void_p(tg.data_ptr()), ctypes.c_void_p(hand.data_ptr()), 1))
    o.dof[2, 0, :23] = 0.3; o.dof[2, 1, :23] = -0.2; o.dof[2, 2, :23] = 0.25
    torch.cuda.synchronize()
    assert np.array_equal(g.tensor("BRICK").cpu().numpy(), o.brick)
    assert np.array_equal(g.tensor("DOF").cpu().numpy(), o.dof)
    for _ in range(2):       # and the simulation continues identically from the written state
        g.simulate(); o.simulate()
    torch.cuda.synchronize()
    assert np.array_equal(g.tensor("BRICK").cpu().numpy(), o.brick)


def test_step_host_matches_device_step(scene, oracle_lib):
    """sdx_step_host (pinned host buffers, VR:165-177 clamps) == device step + clamp"""
    from seqdex_b200.env import SdxEnv
    from tests.util import lattice_bank
    n = 8
    a, b = SdxEnv(scene, n), SdxEnv(scene, n)
    bank = lattice_bank(scene, 2)
    for e in (a, b):
        e.set_heap_bank(bank); e.set_tvalue_weights(oracle_lib.default_tvalue_weights(1))
    act = (torch.rand(n, 23) * 3 - 1.5)
    ha = act.clone().pin_memory()
    obs, st = torch.empty(n, 396).pin_memory(), torch.empty(n, 564).pin_memory()
    rew, rs = torch.empty(n).pin_memory(), torch.empty(n, dtype=torch.int64).pin_memory()
    for _ in range(3):
        a.step_host(ha, obs, st, rew, rs)
        b.step(act.cuda())
    torch.cuda.synchronize()
    assert torch.equal(obs, b.tensor("OBS").clamp(-5, 5).cpu()) and torch.equal(st, b.tensor("STATES").clamp(-5, 5).cpu())
    assert torch.equal(rew, b.tensor("REW").cpu()) and torch.equal(rs, b.tensor("RESET").cpu())


def test_create_fails_loudly_without_bank(scene):
    from seqdex_b200.env import SdxEnv
    e = SdxEnv(scene, 8)
    with pytest.raises(RuntimeError, match="heap bank"):
        e.step(torch.zeros(8, 23, device="cuda"))


def test_facade_dump_is_current():
    """tests/golden/facade_dump.npz -- the facade tensors the reference's own compute_observations is run on in
    tests/test_facade_reference_cpu.py -- is what tools/dump_facade.py produces with today's kernels, bit for bit"""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("dump_facade", os.path.join(here, "..", "tools", "dump_facade.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    new = mod.make_dump()
    old = dict(np.load(os.path.join(here, "golden", "facade_dump.npz")))
    assert set(new) == set(old)
    for k in new:
        assert np.array_equal(new[k], old[k]), f"{k} differs: regenerate with `python tools/dump_facade.py` on a B200 and re-run tests/test_facade_reference_cpu.py"
