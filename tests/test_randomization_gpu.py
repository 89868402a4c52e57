"""Domain randomisation on the GPU: sdx_dr_randn / sdx_dr_noise through the C-ABI against the numpy oracle, and
the task-level hooks (BT:130-150) on BlockAssemblyGraspSim."""
import copy
import ctypes

import numpy as np
import pytest
import torch

from oracle import dr_oracle
from tests.util import lattice_bank

pytestmark = pytest.mark.gpu

CONFIGS = [
    {"range": [0, .002], "range_correlated": [0, .001], "operation": "additive", "distribution": "gaussian", "schedule": "linear", "schedule_steps": 40000},
    {"range": [1.0, .05], "range_correlated": [1.0, .02], "operation": "scaling", "distribution": "gaussian", "schedule": "linear", "schedule_steps": 3000},
    {"range": [-0.01, .02], "range_correlated": [-0.005, .005], "operation": "additive", "distribution": "uniform", "schedule": "linear", "schedule_steps": 1000},
    {"range": [0.9, 1.1], "range_correlated": [0.95, 1.05], "operation": "scaling", "distribution": "uniform"},
]


@pytest.fixture(scope="module")
def env():
    from seqdex_b200.env import SdxEnv
    from seqdex_b200.scene import Scene
    e = SdxEnv(Scene(), 8)
    yield e
    e.close()


def _randn(env, n, seed, counter):
    out = torch.empty(n, device="cuda")
    assert env.L.sdx_dr_randn(env.h, ctypes.c_void_p(out.data_ptr()), ctypes.c_int64(n), ctypes.c_uint64(seed), ctypes.c_uint32(counter)) == 0
    return out


def _noise(env, src, corr, p, seed, counter):
    out = torch.empty_like(src)
    rc = env.L.sdx_dr_noise(env.h, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(corr.data_ptr()),
                            ctypes.c_int64(src.numel()), ctypes.c_float(p["a_corr"]), ctypes.c_float(p["b_corr"]), ctypes.c_float(p["a"]),
                            ctypes.c_float(p["b"]), ctypes.c_int(p["distribution"]), ctypes.c_int(p["operation"]), ctypes.c_uint64(seed),
                            ctypes.c_uint32(counter))
    assert rc == 0
    return out


@pytest.mark.parametrize("n", [1, 3, 4, 1001, 8 * 396])
def test_randn_matches_the_oracle(env, n):
    """same Philox words, Box-Muller in fp32: CUDA's logf / sinf / cosf against numpy's, a few ulp of |z| <= 5.6"""
    z = _randn(env, n, 0x1234567890ABCDEF, 7).cpu().numpy()
    np.testing.assert_allclose(z, dr_oracle.randn(n, 0x1234567890ABCDEF, 7), rtol=0, atol=1e-5)
    assert not np.array_equal(z, _randn(env, n, 0x1234567890ABCDEF, 8).cpu().numpy()) or n == 0


@pytest.mark.parametrize("ci", range(len(CONFIGS)))
def test_noise_matches_the_oracle(env, ci):
    p = dr_oracle.nonphysical_params(dict(CONFIGS[ci]), 700)
    g = torch.Generator(device="cuda").manual_seed(ci)
    src = torch.randn(37, 23, device="cuda", generator=g)
    corr = _randn(env, src.numel(), 99, 1)
    out = _noise(env, src, corr, p, 99, 2).cpu().numpy()
    ref = dr_oracle.noise(src.cpu().numpy(), corr.cpu().numpy(), p, 99, 2)
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-5 * max(1.0, abs(p["a"])) * 8)
    # the deterministic part (no white noise: a = 0) is plain fp32 arithmetic and must agree bit for bit
    q = dict(p, a=0.0)
    out0 = _noise(env, src, corr, q, 99, 3).cpu().numpy()
    ref0 = dr_oracle.combine(src.cpu().numpy().reshape(-1), corr.cpu().numpy(), np.zeros(src.numel(), np.float32), q).reshape(src.shape)
    assert np.array_equal(out0, ref0)
    assert np.array_equal(out, _noise(env, src, corr, p, 99, 2).cpu().numpy())      # same counter, same noise
    assert env.L.sdx_dr_noise(env.h, None, None, None, ctypes.c_int64(4), ctypes.c_float(0), ctypes.c_float(0), ctypes.c_float(0),
                              ctypes.c_float(0), 0, 0, ctypes.c_uint64(0), ctypes.c_uint32(0)) != 0       # bad arguments are refused


def _cfg(n, randomize, params=None):
    from seqdex_b200.tasks.block_assembly_grasp_sim import DEFAULT_CFG
    cfg = copy.deepcopy(DEFAULT_CFG)
    cfg["env"]["numEnvs"] = n
    cfg["task"] = {"randomize": randomize, "randomization_params": params or {}}
    return cfg


def test_observation_noise_does_not_leak_into_the_env(env):
    """BT:149-150 replaces obs_buf by a noisy COPY; the history frames GS:1330-1332 keeps are noise-free.  With observation noise
    only, the env of a randomised task must evolve bit-identically to a plain one, and obs_buf = OBS + noise of the configured size."""
    from seqdex_b200.scene import Scene
    from seqdex_b200.tasks.block_assembly_grasp_sim import BlockAssemblyGraspSim
    from seqdex_b200.vec_task import RLgamesVecTaskPython
    n = 16
    bank = lattice_bank(Scene(), 2)
    params = {"frequency": 2, "observations": {"range": [0, .02], "range_correlated": [0, .01], "operation": "additive",
                                               "distribution": "gaussian"}, "actor_params": {}}
    plain = BlockAssemblyGraspSim(_cfg(n, False), heap_bank=bank)
    noisy = BlockAssemblyGraspSim(_cfg(n, True, params), heap_bank=bank)
    vt = RLgamesVecTaskPython(noisy, "cuda:0")
    g = torch.Generator(device="cuda").manual_seed(5)
    for step in range(5):
        a = torch.rand(n, 23, device="cuda", generator=g) * 2 - 1
        plain.step(a.clone())
        od, rew, reset, _ = vt.step(a.clone())
        assert torch.equal(plain.env.tensor("OBS"), noisy.env.tensor("OBS"))
        assert torch.equal(plain.env.tensor("BRICK"), noisy.env.tensor("BRICK")) and torch.equal(plain.rew_buf, rew)
        d = (noisy.obs_buf - noisy.env.tensor("OBS")).flatten()
        assert 0.015 < float(d.std()) < 0.03 and abs(float(d.mean())) < 0.003     # sqrt(.02^2 + .01^2) = .0224
        assert torch.equal(od["obs"], torch.clamp(noisy.obs_buf, -5, 5))
    r = noisy.randomizer
    assert r.frame == 5 and r.sched.last_rand_step == 4                            # refreshed at frames 0, 2, 4
    obs_out, st_out = torch.empty_like(noisy.obs_buf), torch.empty_like(noisy.states_buf)
    vt.step_into(torch.zeros(n, 23, device="cuda"), obs_out, st_out)
    assert torch.equal(obs_out, torch.clamp(noisy.obs_buf, -5, 5)) and not torch.equal(obs_out, torch.clamp(noisy.env.tensor("OBS"), -5, 5))


def test_action_noise_and_gravity(env):
    from seqdex_b200.scene import Scene
    from seqdex_b200.tasks.block_assembly_grasp_sim import BlockAssemblyGraspSim
    n = 16
    params = {"frequency": 1000,
              "actions": {"range": [0., .05], "range_correlated": [0, .015], "operation": "additive", "distribution": "gaussian"},
              "sim_params": {"gravity": {"range": [0, 0.4], "operation": "additive", "distribution": "gaussian"}}, "actor_params": {}}
    t = BlockAssemblyGraspSim(_cfg(n, True, params), heap_bank=lattice_bank(Scene(), 2))
    r = t.randomizer
    assert r.gravity != r.gravity0 and abs(r.gravity - r.gravity0) < 2.5
    a = torch.zeros(n, 23, device="cuda")
    t.step(a)
    d = t._act_noisy.flatten()                                                   # what pre_physics_step received
    assert 0.03 < float(d.std()) < 0.08
    assert torch.isfinite(t.obs_buf).all() and torch.isfinite(t.env.tensor("BRICK")).all()
    with pytest.raises(NotImplementedError):
        BlockAssemblyGraspSim(_cfg(n, True, {"actor_params": {"lego": {"rigid_body_properties": {"mass": {"range": [0.5, 1.5]}}}}}),
                              heap_bank=lattice_bank(Scene(), 2))
