"""Rows a13-a16 on the GPU against vectors produced by EXECUTING the reference's own Python (oracle/gen_golden_ppo.py ->
tests/golden/ppo_*.npz, tvalue_trainer.npz), through the C-ABI.  fp32 kernels: 1e-5 relative; the tensor-core t-value step: bf16."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import ppo_oracle as PO

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
C = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
P = lambda t: ctypes.c_void_p(t.data_ptr())
ST = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def test_actor_loss_kernel_reproduces_the_reference_neglogp():
    """old_neglogp := the reference's _calc_neglogp (RGC:2113-2127) of the same (x, mean, logstd): the kernel's own neglogp must
    give ratio = exp(old - new) = 1 for every sample, i.e. a_loss = -adv exactly and the unclipped gradient"""
    from seqdex_b200 import _lib
    L = _lib.load()
    g = np.load(os.path.join(G, "ppo_neglogp.npz"))
    x, mean, logstd, nlp = C(g["x"]), C(g["mean"]), C(g["logstd"]), C(g["neglogp"])
    M, A = x.shape
    adv = torch.linspace(-1, 1, M, device="cuda").contiguous()
    dmu, dls, stats = torch.zeros(M, A, device="cuda"), torch.zeros(A, device="cuda"), torch.zeros(4, device="cuda")
    _lib.check(L.sdx_ppo_actor_loss(P(mean), P(logstd), P(x), P(mean), P(logstd), P(nlp), P(adv), M, A, ctypes.c_float(0.1), ctypes.c_float(0.0),
                                    ctypes.c_float(1.0), P(dmu), P(dls), P(stats), ST()))
    torch.cuda.synchronize()
    assert abs(float(stats[0]) + float(adv.sum())) < 1e-3                 # sum of -adv * 1
    ref = adv[:, None] * (-(x - mean) / torch.exp(logstd) ** 2)            # d(-adv * ratio)/d mu at ratio = 1
    torch.testing.assert_close(dmu, ref, rtol=2e-4, atol=2e-5)
    assert abs(float(stats[2])) < 1e-2                                     # KL(new || old) of identical distributions: M x A x log(1 + 1e-5)


def test_gae_kernel_on_the_arguments_the_reference_passes():
    from seqdex_b200 import _lib
    L = _lib.load()
    g = np.load(os.path.join(G, "ppo_play_steps.npz"))
    H, N = g["rew_stream"].shape[:2]
    rew, val, dn = C(g["gae_mb_rewards"][..., 0]), C(g["gae_mb_values"][..., 0]), C(g["gae_mb_fdones"])
    lv, ld = C(g["gae_last_values"][..., 0]), C(g["gae_fdones"])
    adv, ret = torch.zeros(H, N, device="cuda"), torch.zeros(H, N, device="cuda")
    _lib.check(L.sdx_gae(P(rew), P(val), P(dn), P(lv), P(ld), P(adv), P(ret), H, N, ctypes.c_float(0.99), ctypes.c_float(0.95), ST()))
    torch.cuda.synchronize()
    torch.testing.assert_close(adv.cpu(), torch.from_numpy(g["gae_advs"][..., 0]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(PO.swap_and_flatten01(ret).cpu(), torch.from_numpy(g["batch_returns"][:, 0]), rtol=1e-5, atol=1e-6)


def test_play_steps_stores_what_the_reference_stores():
    """A2CAgent.play_steps on the scripted env / scripted networks of the golden: the rollout buffers hold what the reference's
    experience buffer holds (obs and dones from BEFORE the step), and GAE sees the same arrays (RGC:1394-1483)"""
    from seqdex_b200.ppo import A2CAgent, PPOConfig
    g = np.load(os.path.join(G, "ppo_play_steps.npz"))
    H, N, OD = g["obs_stream"].shape[0] - 1, g["obs_stream"].shape[1], g["obs_stream"].shape[2]
    SD = g["st_stream"].shape[2]
    obs_s, st_s, rew_s, done_s, val_s = C(g["obs_stream"]), C(g["st_stream"]), C(g["rew_stream"]), C(g["done_stream"]), C(g["val_stream"])

    class Env:
        num_envs, num_actions, num_obs, num_states = N, 23, OD, SD
        t = 0

        def step(self, a):
            t = self.t
            self.t += 1
            return {"obs": obs_s[t + 1], "states": st_s[t + 1]}, rew_s[t, :, 0], done_s[t].float(), {}

    agent = A2CAgent(Env(), PPOConfig(minibatch_size=H * N, cv_normalize_input=False))
    agent.set_obs(obs_s[0], st_s[0])
    agent.dones.copy_(C(g["dones0"]).float())
    k = {"cv": 0}

    def cv_forward(states, mean=None, var=None, train=False):
        t = k["cv"]
        k["cv"] += 1
        assert torch.equal(states, st_s[t])
        return val_s[t]
    agent.cv.forward = cv_forward
    agent.play_steps()
    torch.cuda.synchronize()
    assert torch.equal(agent.b_obs, obs_s[:H]) and torch.equal(agent.b_states, st_s[:H])
    np.testing.assert_array_equal(agent.b_dones.cpu().numpy(), g["gae_mb_fdones"])
    np.testing.assert_array_equal(agent.b_values.cpu().numpy(), g["gae_mb_values"][..., 0])
    np.testing.assert_array_equal(agent.b_rewards.cpu().numpy(), g["gae_mb_rewards"][..., 0])
    np.testing.assert_array_equal(agent.dones.cpu().numpy(), g["gae_fdones"])
    np.testing.assert_array_equal(agent.last_values.cpu().numpy(), g["gae_last_values"][..., 0])
    torch.testing.assert_close(agent.b_adv.cpu(), torch.from_numpy(g["gae_advs"][..., 0]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(PO.swap_and_flatten01(agent.b_returns).cpu(), torch.from_numpy(g["batch_returns"][:, 0]), rtol=1e-5, atol=1e-6)


def test_env_major_batch_conversion():
    """the bf16 batch the update slices its minibatches from is rl_games' env-major swap_and_flatten01 of the time-major buffers"""
    from seqdex_b200.ppo import MLP
    H, N, D = 8, 24, 70
    m = MLP(D, 3, H * N, seed=1)
    x = torch.randn(H, N, D, device="cuda")
    xb = torch.zeros(H * N, m.in_pad, device="cuda", dtype=torch.bfloat16)
    xt = torch.zeros(m.in_pad + 16, H * N, device="cuda", dtype=torch.bfloat16)
    m.convert_batch_env_major(x.view(H * N, D), H, xb, xt)
    torch.cuda.synchronize()
    ref = PO.swap_and_flatten01(x).to(torch.bfloat16)
    assert torch.equal(xb[:, :D], ref) and float(xb[:, D:].float().abs().max()) == 0.0
    assert torch.equal(xt[:D], ref.T) and float(xt[m.in_pad].float().min()) == 1.0


def test_advantage_normalisation_kernels_vs_prepare_dataset():
    from seqdex_b200 import _lib
    L = _lib.load()
    g = np.load(os.path.join(G, "ppo_prepare_dataset.npz"))
    adv = C((g["returns"] - g["values"])[:, 0])
    mom = torch.zeros(2, device="cuda", dtype=torch.float64)
    _lib.check(L.sdx_moments(P(adv), adv.numel(), P(mom), ST()))
    _lib.check(L.sdx_normalize(P(adv), adv.numel(), P(mom), ctypes.c_double(adv.numel()), ST()))
    torch.cuda.synchronize()
    torch.testing.assert_close(adv.cpu(), torch.from_numpy(g["advantages"]), rtol=1e-5, atol=1e-6)


def test_device_side_adaptive_lr_follows_the_legacy_schedule():
    """sdx_ppo_adaptive_lr after every minibatch == the lr sequence the reference's train_epoch produces (RGC:1360-1365)"""
    from seqdex_b200 import _lib
    L = _lib.load()
    g = np.load(os.path.join(G, "ppo_schedule_legacy.npz"))
    mb = 4096
    lr, acc, stats = torch.full((1,), float(g["lr0"]), device="cuda"), torch.zeros(8, device="cuda"), torch.zeros(4, device="cuda")
    out = []
    for kl in g["kls"]:
        stats[2] = float(kl) * mb
        stats[0] = 1.0
        _lib.check(L.sdx_ppo_adaptive_lr(P(stats), ctypes.c_float(1.0 / mb), ctypes.c_float(float(g["kl_threshold"])), ctypes.c_float(1e-6),
                                         ctypes.c_float(1e-2), P(lr), P(acc), 1, ST()))
        out.append(float(lr))
    np.testing.assert_allclose(out, g["lrs"], rtol=1e-5)
    assert float(acc[4]) == len(g["kls"]) and float(stats.abs().max()) == 0.0 and abs(float(acc[0]) - len(g["kls"])) < 1e-6
    np.testing.assert_allclose(float(acc[5]), float(g["kls"][-1]), rtol=1e-5)


def test_tvalue_trainer_step_vs_the_reference_trainer():
    """one TValue_Trainer.train_rollout iteration (TVT:207-229) executed by the reference vs ours on the same rows, noise and
    initial weights: batch construction fp32-exact; logits / loss within bf16 tensor-core tolerance; the Adam step moves the
    weights the same way"""
    from seqdex_b200.tvalue import TValueTrainer
    g = np.load(os.path.join(G, "tvalue_trainer.npz"))
    tr = TValueTrainer(g["success_data"], g["failure_data"], seed=0)
    tr.net.load_flat(torch.from_numpy(g["w0"]))
    succ, fail, rf = C(g["success_data"]), C(g["failure_data"]), C(g["rand_float"])
    x, y = tr.make_batch(succ[C(g["succ_rand"])], fail[C(g["fail_rand"])], rf[:, 0:4], rf[:, 4:8])
    torch.testing.assert_close(x.cpu(), torch.from_numpy(g["obs_buf"]), rtol=1e-6, atol=1e-7)
    assert torch.equal(y.cpu().long(), torch.from_numpy(g["target"][:, 1]).long())
    z = torch.nn.functional.elu(tr.net.forward(x, train=True).clone())
    assert float((z.cpu() - torch.from_numpy(g["logits"])).abs().max()) < 2e-2
    stats = tr.step(x, y)
    torch.cuda.synchronize()
    assert abs(float(stats[0]) / (2 * x.shape[0]) - float(g["loss"])) < 2e-3
    dw, dref = tr.net.params.cpu() - torch.from_numpy(g["w0"]), torch.from_numpy(g["w1"] - g["w0"])
    big = dref.abs() > 9e-4                                              # first Adam step: |dw| = lr wherever the gradient is not ~0
    assert float((torch.sign(dw[big]) == torch.sign(dref[big])).float().mean()) > 0.97
    assert float((dw - dref).abs().mean()) < 2e-4
