"""SURVEY.md 8f.2 on the GPU: rl_games checkpoints through a live agent, the frozen inner policy (``NNController``), and
the two-agent ``PolicySequencingRunner`` (PSR:39-373) on one env."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = {"env": {"numEnvs": 256, "episodeLength": 150, "actionsMovingAverage": 1.0}, "sim": {"substeps": 2, "physx": {}}, "task": {"randomize": False}}


def _agent(scene, n=256, mb=1024, seed=22):
    from seqdex_b200.ppo import A2CAgent, PPOConfig
    from seqdex_b200.tasks import BlockAssemblyGraspSim
    from seqdex_b200.vec_task import RLgamesVecTaskPython
    from tests.util import lattice_bank
    cfg = {**CFG, "env": {**CFG["env"], "numEnvs": n}}
    task = BlockAssemblyGraspSim(cfg, heap_bank=lattice_bank(scene, 2))
    return A2CAgent(RLgamesVecTaskPython(task, "cuda:0"), PPOConfig(minibatch_size=mb, seed=seed))


def test_agent_checkpoint_round_trip_resumes_bit_exactly(scene, tmp_path):
    """save -> restore into a fresh agent -> both continue with identical updates (weights, Adam moments, step counters,
    RunningMeanStd, lr): what ``--checkpoint`` / ``_restore`` promise (PSR:74-75)."""
    a = _agent(scene)
    a.train_epoch()
    fn = a.save(os.path.join(tmp_path, "nn", "last_allegro_ep_1"))
    ck = torch.load(fn, map_location="cpu", weights_only=False)
    assert set(ck) >= {"model", "epoch", "optimizer", "assymetric_vf_nets", "frame", "last_mean_rewards", "env_state"}
    assert sum(v.numel() for v in ck["model"].values()) == 2_131_503
    b = _agent(scene, seed=5)                       # different initial weights: everything must come from the file
    b.restore(fn)
    assert torch.equal(a.actor.params, b.actor.params) and torch.equal(a.cv.params, b.cv.params)
    assert torch.equal(a.actor.adam_m, b.actor.adam_m) and torch.equal(a.cv.adam_v, b.cv.adam_v)
    assert torch.equal(a.rms_mean, b.rms_mean) and torch.equal(a.rms_var, b.rms_var) and torch.equal(a.rms_count, b.rms_count)
    assert a._adam_step(a.actor) == b._adam_step(b.actor) > 0 and a.last_lr == b.last_lr and b.epoch_num == 1
    # one identical gradient step on both: parameters stay bit-identical
    g = torch.randn_like(a.actor.grads) * 1e-3
    for ag in (a, b):
        ag.actor.grads.copy_(g)
        ag.actor.adam(ag.last_lr, 1.0)
    torch.cuda.synchronize()
    assert torch.equal(a.actor.params, b.actor.params)


def test_nn_controller_predict_matches_torch(tmp_path):
    """nn_controller.py:27-58: load ``checkpoint['model']``, deterministic predict = clip(mu), stochastic = clip(mu + eps)"""
    from seqdex_b200 import checkpoint as ck
    from seqdex_b200.policy_sequencing import NNController
    units, obs_dim = (512, 256, 128), 81                   # utils/robot_controller/network.yaml
    n = ck.mlp_slices(obs_dim, 23, units, has_sigma=True)[1]
    flat = torch.randn(n, generator=torch.Generator().manual_seed(0)) * 0.05
    flat[n - 23:] = -1.0                                    # log sigma
    fn = ck.save_checkpoint(os.path.join(tmp_path, "policy"), {"model": ck.actor_state_dict(flat, obs_dim, 23, units)})
    pol = NNController(num_actors=8, units=units, obs_dim=obs_dim)
    pol.load(fn)
    obs = torch.randn(8, obs_dim, device="cuda").clamp(-5, 5)
    act = pol.predict(obs, deterministic=True)
    ref = pol.model.torch_reference()(obs).clamp(-1, 1)
    assert act.shape == (8, 23) and float((act - ref).abs().max()) < 2e-2
    s1, s2 = pol.predict(obs), pol.predict(obs)
    assert float((s1 - s2).abs().max()) > 0                 # fresh noise each call
    assert float(s1.abs().max()) <= 1.0
    noise = (s1 - ref)[(s1.abs() < 1) & (ref.abs() < 1)]
    assert abs(float(noise.std()) - math.exp(-1.0)) < 0.1   # sigma = exp(logstd)
    single = pol.predict(obs[0].cpu().numpy(), deterministic=True)    # numpy, one observation (nn_controller.py:28-29)
    assert single.shape == (1, 23) and float((single[0] - act[0]).abs().max()) < 1e-6


class _SequencedGraspSim:
    """a sequenced task in the sense of vec_task_lego.py: GraspSim publishes its buffers as the 'before' phase (the learned
    grasp, progress < 100) and as the 'after' phase (progress >= 100, the scripted insert preparation GS:1625-1634)."""

    def __init__(self, task):
        self.t = task
        for k in ("num_envs", "num_obs", "num_states", "num_actions", "device", "progress_buf", "rew_buf", "reset_buf"):
            setattr(self, k, getattr(task, k))
        self.grasping_num_obs = self.insertion_num_obs = task.num_obs
        self.grasping_num_states = self.insertion_num_states = task.num_states
        self.extras = {}

    def step(self, actions):
        self.t.step(actions)
        t = self.t
        self.extras = {"before_obs": t.obs_buf, "before_states": t.states_buf, "after_obs": t.obs_buf, "after_states": t.states_buf,
                       "before_rew_buf": t.rew_buf, "after_rew_buf": t.rew_buf, "before_reset_buf": t.reset_buf, "after_reset_buf": t.reset_buf}


def test_policy_sequencing_runner_hands_over_at_before_episode_length(scene):
    from seqdex_b200.policy_sequencing import LegoVecTaskPython, PolicySequencingRunner
    from seqdex_b200.ppo import PPOConfig
    from seqdex_b200.tasks import BlockAssemblyGraspSim
    from tests.util import lattice_bank
    task = BlockAssemblyGraspSim(CFG, heap_bank=lattice_bank(scene, 2))
    env = LegoVecTaskPython(_SequencedGraspSim(task), "cuda:0")
    r = PolicySequencingRunner(env, PPOConfig(minibatch_size=1024), PPOConfig(minibatch_size=1024, seed=23), before_episode_length=12)
    p = [ag.actor.params.clone() for ag in r.agents]
    out = r.run(1)                                           # steps 1..8: progress_buf[0] < 12 -> 'before' acts and learns
    assert out[0]["trained"] == "before" and all(math.isfinite(v) for v in out[0].values() if isinstance(v, float))
    assert float((r.agents[0].actor.params - p[0]).abs().max()) > 0 and torch.equal(r.agents[1].actor.params, p[1])
    out = r.run(1)                                           # steps 9..16: the hand-over happens inside this rollout
    assert out[0]["trained"] == "after"
    assert float((r.agents[1].actor.params - p[1]).abs().max()) > 0
    assert int(task.progress_buf[0]) == 17                   # reset() step + 16 sequenced steps, nobody reset in between
