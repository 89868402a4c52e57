"""Domain randomisation, CPU side: the oracle and the host-side parameter arithmetic against golden vectors
produced by executing the reference's BaseTask.apply_randomizations and its noise_lambda closures (oracle/gen_golden_dr.py)."""
import json
import os

import numpy as np
import pytest

from oracle import dr_oracle
from seqdex_b200 import randomization as R

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(G, "dr_params.npz")), json.load(open(os.path.join(G, "dr_configs.json")))


@pytest.mark.parametrize("impl", [dr_oracle.nonphysical_params, R.nonphysical_params])
def test_schedule_parameters_match_the_reference(golden, impl):
    g, configs = golden
    for ci, cfg in enumerate(configs):
        keys = ("mu", "var", "mu_corr", "var_corr") if cfg["distribution"] == "gaussian" else ("lo", "hi", "lo_corr", "hi_corr")
        for si, step in enumerate(g["steps"]):
            p = impl(dict(cfg), int(step))
            np.testing.assert_allclose([p[k] for k in keys], g[f"c{ci}_params"][si], rtol=0, atol=0)   # python floats: exact


def test_noise_lambda_formula_matches_the_reference(golden):
    """the closure's arithmetic, fed the very draws torch made: fp32, within 1 ulp of torch's fused elementwise kernels"""
    g, configs = golden
    n = 0
    for ci, cfg in enumerate(configs):
        for step in (250, 2500, 123456):
            tag = f"c{ci}_s{step}_"
            p = dr_oracle.nonphysical_params(dict(cfg), step)
            for w, y in (("w1", "y1"), ("w2", "y2")):
                out = dr_oracle.combine(g[tag + "x"], g[tag + "corr"], g[tag + w], p)
                np.testing.assert_allclose(out, g[tag + y], rtol=3e-7, atol=1e-9)
                n += 1
    assert n == 42


def test_refresh_bookkeeping_matches_the_reference(golden):
    """first call randomises; afterwards the non-env parameters are regenerated when frequency frames have passed (BT:233-249)"""
    g, configs = golden
    sched = R.RefreshSchedule(int(g["refresh_freq"][0]))
    for frame, refreshed, last in g["refresh_log"]:
        assert int(sched.due(int(frame))) == int(refreshed)
        assert sched.last_rand_step == int(last)


def test_white_noise_stream_statistics():
    z = dr_oracle.randn(1 << 18, 22, 3)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01 and np.isfinite(z).all()
    u = dr_oracle.white(1 << 18, 22, 3, True)
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    assert not np.array_equal(z, dr_oracle.randn(1 << 18, 22, 4))           # the call counter selects a fresh stream
    assert np.array_equal(z[:1001], dr_oracle.randn(1001, 22, 3))           # element i does not depend on the length


def test_physical_sample_restatement():
    """generate_random_samples (isaacgym.gymutil, restated): schedule interpolation and ranges"""
    rng = np.random.default_rng(0)
    cfg = {"range": [0.5, 1.5], "operation": "scaling", "distribution": "uniform", "schedule": "linear", "schedule_steps": 1000}
    assert np.allclose(dr_oracle.generate_random_samples(cfg, 8, 0, rng), 1.0)                   # no randomisation at step 0
    s = dr_oracle.generate_random_samples(cfg, 4096, 500, rng)
    assert 0.75 <= s.min() and s.max() <= 1.25
    cfg = {"range": [0, 0.4], "operation": "additive", "distribution": "gaussian", "schedule": "linear", "schedule_steps": 40000}
    s = dr_oracle.generate_random_samples(cfg, 20000, 40000, rng)
    assert abs(s.std() - 0.4) < 0.02 and abs(s.mean()) < 0.02
    cfg = {"range": [0.3, 3.0], "operation": "scaling", "distribution": "loguniform"}
    s = dr_oracle.generate_random_samples(cfg, 4096, 7, rng)
    assert 0.3 <= s.min() and s.max() <= 3.0 and abs(np.log(s).mean() - 0.5 * (np.log(0.3) + np.log(3.0))) < 0.05


def test_unsupported_sections_fail_loudly():
    with pytest.raises(NotImplementedError):
        R.check_supported({"actor_params": {"hand": {"rigid_body_properties": {"mass": {}}}}})
    R.check_supported({"frequency": 10, "observations": {}, "actions": {}, "sim_params": {"gravity": {}}, "actor_params": {}})
    with pytest.raises(NotImplementedError):
        R.check_supported({"sim_params": {"dt": {}}})


class _RecordingLib:
    """stands in for the CUDA library: records the C-ABI calls DomainRandomizer makes (host logic only, no compute)"""

    def __init__(self):
        self.calls = []

    def _rec(self, name):
        def f(*args):
            self.calls.append((name, [getattr(a, "value", a) for a in args]))
            return 0
        return f

    def __getattr__(self, name):
        if name.startswith("sdx_"):
            return self._rec(name)
        raise AttributeError(name)


def test_domain_randomizer_host_logic():
    """the call sequence of a randomised task over a refresh (BT:229-340): parameters from the schedule at the current frame, the
    correlated tensor redrawn exactly once per refresh, a fresh Philox counter for every kernel call, gravity set on refresh only"""
    import types
    import torch
    lib = _RecordingLib()
    scene = types.SimpleNamespace(c=types.SimpleNamespace(gravity_z=-9.81))
    env = types.SimpleNamespace(L=lib, h=1234, n=4, device=torch.device("cpu"), scene=scene)
    params = {"frequency": 3,
              "observations": {"range": [0, .002], "range_correlated": [0, .001], "operation": "additive", "distribution": "gaussian",
                               "schedule": "linear", "schedule_steps": 10},
              "sim_params": {"gravity": {"range": [0, 0.4], "operation": "additive", "distribution": "gaussian"}},
              "actor_params": {}}
    r = R.DomainRandomizer(env, params, seed=22)
    obs, out = torch.zeros(4, 6), torch.zeros(4, 6)
    acts = torch.ones(4, 23)
    refreshed = []
    for step in range(7):
        refreshed.append(r.apply_randomizations(torch.ones(4, dtype=torch.int64)))
        assert r.noise("actions", acts, None) is acts                      # not configured: the tensor itself comes back
        assert r.noise("observations", obs, out) is out
        r.step_done()
    assert refreshed == [True, False, False, True, False, False, True]    # frames 0, 3, 6
    names = [c[0] for c in lib.calls]
    assert names.count("sdx_set_gravity") == 3 and names.count("sdx_dr_randn") == 3 and names.count("sdx_dr_noise") == 7
    counters = [c[1][-1] for c in lib.calls if c[0] in ("sdx_dr_randn", "sdx_dr_noise")]
    assert counters == sorted(set(counters)) and len(counters) == 10      # every kernel call gets its own stream
    seeds = {c[1][-2] for c in lib.calls if c[0] in ("sdx_dr_randn", "sdx_dr_noise")}
    assert len(seeds) == 1
    # the numbers handed to sdx_dr_noise are the schedule's at the frame of the last refresh (frame 3 -> s = 0.3, frame 6 -> 0.6)
    noise_calls = [c[1] for c in lib.calls if c[0] == "sdx_dr_noise"]
    for k, frame in ((3, 3), (4, 3), (5, 3), (6, 6)):
        p = R.nonphysical_params(dict(params["observations"]), frame)
        a_corr, b_corr, a, b, dist, op = noise_calls[k][5:11]
        assert (dist, op) == (0, 0)
        np.testing.assert_allclose([a_corr, b_corr, a, b], [p["a_corr"], p["b_corr"], p["a"], p["b"]], rtol=1e-6, atol=0)
    assert noise_calls[0][4] == 24                                        # element count
    assert r.frame == 7 and int(r.randomize_buf[0]) == 1                  # BT:246-249: zeroed at frames 3 and 6 (due AND resetting), +1 per step
