"""PPO tensor path (csrc/sdx_ppo.cu via seqdex_b200.ppo) against plain PyTorch fp32 references of the same ops
(rl_games 1.5.2 formulas).  bf16 tensor-core compute -> tolerances: forward 2e-2 relative to the
output scale, gradients: relative Frobenius error < 3e-2 and cosine > 0.999; fp32 elementwise kernels: 1e-5."""
import ctypes
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _relerr(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("in_dim,out_dim,M", [(396, 23, 1024), (564, 1, 640)])
def test_mlp_forward_backward_vs_torch(in_dim, out_dim, M):
    from seqdex_b200.ppo import MLP
    torch.manual_seed(0)
    m = MLP(in_dim, out_dim, 2048, has_sigma=(out_dim > 1), seed=3)
    ref = m.torch_reference()
    x = torch.randn(M, in_dim, device="cuda").clamp(-5, 5)
    out = m.forward(x, train=True).clone()
    xr = x.clone().requires_grad_(False)
    out_ref = ref(xr)
    scale = float(out_ref.abs().mean())
    assert float((out - out_ref).abs().max()) < 3e-2 * max(scale, 1.0), (float((out - out_ref).abs().max()), scale)
    dout = torch.randn(M, out_dim, device="cuda") / M
    m.backward(dout.contiguous())
    torch.cuda.synchronize()
    out_ref.backward(dout)
    gref = torch.cat([p.grad.reshape(-1) for lin in ref if isinstance(lin, torch.nn.Linear) for p in (lin.weight, lin.bias)])
    n = gref.numel()
    g = m.grads[:n]
    assert torch.isfinite(g).all()
    cos = float(torch.nn.functional.cosine_similarity(g, gref, dim=0))
    assert cos > 0.999 and _relerr(g, gref) < 3e-2, (cos, _relerr(g, gref))
    for name, off, shp in m.slices():     # every layer individually (a wrong bias column would hide in the global norm)
        if name == "sigma":
            continue
        k = int(torch.tensor(shp).prod())
        assert _relerr(g[off:off + k], gref[off:off + k]) < 6e-2, (name, _relerr(g[off:off + k], gref[off:off + k]))


def test_mlp_input_normalisation():
    from seqdex_b200.ppo import MLP
    torch.manual_seed(1)
    m = MLP(564, 1, 512, seed=5)
    x = torch.randn(256, 564, device="cuda") * 3 + 1
    mean, var = torch.randn(564, device="cuda"), torch.rand(564, device="cuda") + 0.5
    out = m.forward(x, mean, var).clone()
    xn = ((x - mean) / torch.sqrt(var + 1e-5)).clamp(-5, 5)
    ref = m.torch_reference()(xn)
    assert float((out - ref).abs().max()) < 3e-2 * max(float(ref.abs().mean()), 1.0)


def test_adam_matches_torch():
    from seqdex_b200.ppo import MLP
    m = MLP(396, 23, 256, has_sigma=True, seed=7)
    p0 = m.params.clone()
    tp = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([tp], lr=3e-4, eps=1e-8)
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(3):
        grad = torch.randn(m.nparams, device="cuda", generator=g) * 0.01
        m.grads.copy_(grad)
        m.adam(3e-4, max_norm=1.0)
        tp.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([tp], 1.0)
        opt.step()
    torch.cuda.synchronize()
    torch.testing.assert_close(m.params, tp.data, rtol=1e-5, atol=1e-7)


def test_actor_loss_kernel_vs_torch():
    from seqdex_b200 import _lib
    L = _lib.load()
    torch.manual_seed(2)
    M, A, eclip, bc = 1000, 23, 0.1, 0.001
    mu = (torch.randn(M, A, device="cuda") * 0.8).requires_grad_(True)
    logstd = (torch.randn(A, device="cuda") * 0.1).requires_grad_(True)
    old_mu = mu.detach() + torch.randn(M, A, device="cuda") * 0.05
    old_logstd = logstd.detach() + 0.02
    actions = old_mu + torch.exp(old_logstd) * torch.randn(M, A, device="cuda")
    adv = torch.randn(M, device="cuda")

    def neglogp(a, m_, ls):
        return 0.5 * (((a - m_) / torch.exp(ls)) ** 2).sum(-1) + 0.5 * A * math.log(2 * math.pi) + ls.sum()
    old_nlp = neglogp(actions, old_mu, old_logstd)
    nlp = neglogp(actions, mu, logstd)
    ratio = torch.exp(old_nlp - nlp)
    a_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - eclip, 1 + eclip))
    b_loss = (torch.clamp_min(mu - 1.1, 0) ** 2 + torch.clamp_max(mu + 1.1, 0) ** 2).sum(-1)
    loss = a_loss.mean() + bc * b_loss.mean()
    loss.backward()
    sig, so = torch.exp(logstd.detach()), torch.exp(old_logstd)
    from oracle import ppo_oracle as PO
    kl = PO.policy_kl(mu.detach(), sig.expand(M, A), old_mu, so.expand(M, A))      # policy_kl(mu, sigma, old_mu, old_sigma), RGC:1903
    dmu, dls, stats = torch.zeros(M, A, device="cuda"), torch.zeros(A, device="cuda"), torch.zeros(4, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(L.sdx_ppo_actor_loss(p(mu.detach().contiguous()), p(logstd.detach()), p(actions), p(old_mu), p(old_logstd), p(old_nlp), p(adv), M, A,
                                    ctypes.c_float(eclip), ctypes.c_float(bc), ctypes.c_float(1.0 / M), p(dmu), p(dls), p(stats),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    torch.testing.assert_close(dmu, mu.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(dls, logstd.grad, rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(stats[0] / M, a_loss.mean().detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(stats[1] / M, b_loss.mean().detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(stats[2] / M, kl.mean(), rtol=1e-3, atol=1e-6)


def test_value_loss_kernel_vs_torch():
    from seqdex_b200 import _lib
    L = _lib.load()
    torch.manual_seed(3)
    M, eclip = 777, 0.1
    v = torch.randn(M, device="cuda").requires_grad_(True)
    vo, r = v.detach() + torch.randn(M, device="cuda") * 0.15, torch.randn(M, device="cuda")
    vc = vo + (v - vo).clamp(-eclip, eclip)
    loss = torch.max((v - r) ** 2, (vc - r) ** 2).mean()
    loss.backward()
    dv, stats = torch.zeros(M, device="cuda"), torch.zeros(4, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(L.sdx_ppo_value_loss(p(v.detach()), p(vo), p(r), M, ctypes.c_float(eclip), 1, ctypes.c_float(1.0 / M), p(dv), p(stats),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    torch.testing.assert_close(dv, v.grad, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(stats[0] / M, loss.detach(), rtol=1e-5, atol=1e-7)


def test_sampling_and_moments():
    from seqdex_b200 import _lib
    L = _lib.load()
    M, A = 20000, 23
    mu = torch.randn(M, A, device="cuda")
    logstd = torch.full((A,), -0.3, device="cuda")
    act, nlp = torch.zeros(M, A, device="cuda"), torch.zeros(M, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.sdx_ppo_sample(p(mu), p(logstd), M, A, ctypes.c_uint64(5), 11, p(act), p(nlp), st))
    z = (act - mu) / math.exp(-0.3)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    ref = 0.5 * (z ** 2).sum(-1) + 0.5 * A * math.log(2 * math.pi) + float(logstd.sum())
    torch.testing.assert_close(nlp, ref, rtol=1e-5, atol=1e-4)
    x = torch.randn(100000, device="cuda") * 3 + 2
    mom = torch.zeros(2, device="cuda", dtype=torch.float64)
    xr = (x - x.mean()) / (x.std() + 1e-8)
    _lib.check(L.sdx_moments(p(x), x.numel(), p(mom), st))
    _lib.check(L.sdx_normalize(p(x), x.numel(), p(mom), ctypes.c_double(x.numel()), st))
    torch.cuda.synchronize()
    torch.testing.assert_close(x, xr, rtol=1e-4, atol=1e-5)


def test_running_mean_std_merge():
    from seqdex_b200 import _lib
    L = _lib.load()
    D = 564
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    mean, var = torch.zeros(D, device="cuda"), torch.ones(D, device="cuda")
    cnt = torch.full((1,), 1e-4, device="cuda", dtype=torch.float64)
    col = torch.zeros(2 * D, device="cuda", dtype=torch.float64)
    rm, rv, rc = torch.zeros(D, dtype=torch.float64), torch.ones(D, dtype=torch.float64), 1e-4
    for i in range(3):
        x = torch.randn(4096, D, device="cuda") * (i + 1) + i
        _lib.check(L.sdx_col_moments(p(x), 4096, D, p(col), st))
        _lib.check(L.sdx_rms_merge(p(mean), p(var), p(cnt), p(col), D, ctypes.c_double(4096), st))
        xb = x.double().cpu()
        bm, bv, bc = xb.mean(0), xb.var(0), 4096
        delta, tot = bm - rm, rc + bc
        m2 = rv * rc + bv * bc + delta ** 2 * rc * bc / tot
        rm, rv, rc = rm + delta * bc / tot, m2 / tot, tot
    torch.cuda.synchronize()
    torch.testing.assert_close(mean.double().cpu(), rm, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(var.double().cpu(), rv, rtol=1e-4, atol=1e-5)


def test_agent_trains_end_to_end(scene):
    """two PPO iterations on 512 envs through VecTask: finite losses, parameters move, KL small and positive"""
    from seqdex_b200.ppo import A2CAgent, PPOConfig
    from seqdex_b200.tasks import BlockAssemblyGraspSim
    from seqdex_b200.vec_task import RLgamesVecTaskPython
    from tests.util import lattice_bank
    cfg = {"env": {"numEnvs": 512, "episodeLength": 150, "actionsMovingAverage": 1.0}, "sim": {"substeps": 2, "physx": {}}, "task": {"randomize": False}}
    task = BlockAssemblyGraspSim(cfg, heap_bank=lattice_bank(scene, 4))
    env = RLgamesVecTaskPython(task, "cuda:0")
    agent = A2CAgent(env, PPOConfig(minibatch_size=2048))
    p0 = agent.actor.params.clone()
    for _ in range(2):
        info = agent.train_epoch()
        assert all(math.isfinite(v) for v in info.values()), info
    assert float((agent.actor.params - p0).abs().max()) > 0
    assert 0 <= info["kl"] < 0.5
    assert torch.isfinite(agent.b_adv).all() and torch.isfinite(agent.cv.params).all()


def test_preconverted_batch_path_matches_direct_path():
    """forward/backward on a row range of a batch converted once per iteration == forward/backward on that slice"""
    from seqdex_b200.ppo import MLP
    torch.manual_seed(4)
    B, M, row0 = 2048, 512, 1024
    m = MLP(564, 1, 1024, seed=9)
    x = torch.randn(B, 564, device="cuda") * 2
    mean, var = torch.randn(564, device="cuda") * 0.1, torch.rand(564, device="cuda") + 0.5
    out_a = m.forward(x[row0:row0 + M].contiguous(), mean, var, train=True).clone()
    dout = torch.randn(M, 1, device="cuda") / M
    m.backward(dout)
    g_a = m.grads.clone()
    xb = torch.zeros(B, m.in_pad, device="cuda", dtype=torch.bfloat16)
    xt = torch.zeros(m.in_pad + 16, B, device="cuda", dtype=torch.bfloat16)
    m.convert_batch(x, xb, xt, mean, var)
    out_b = m.forward_pre(xb, xt, row0, M).clone()
    m.backward(dout)
    torch.cuda.synchronize()
    assert torch.equal(out_a, out_b)
    torch.testing.assert_close(m.grads, g_a, rtol=1e-4, atol=1e-6)     # split-K atomics: order-dependent last bits
    assert float(xt[m.in_pad].float().min()) == 1.0 and float(xt[m.in_pad + 1:].abs().max()) == 0.0


def test_tvalue_trainer_learns_and_matches_torch_loss(scene):
    """TVT:180-248 semantics: one loss/gradient evaluation against torch (BCEWithLogits on ELU outputs), then training on a
    separable synthetic set reaches > 90 % held-out accuracy and the weights drive the env's gate kernel."""
    import numpy as np
    from seqdex_b200.tvalue import TValueTrainer
    rng = np.random.default_rng(0)
    q = rng.normal(size=(6000, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    succ, fail = q[q[:, 3] > 0.15], q[q[:, 3] < -0.15]          # feasible iff the camera-frame quaternion has w > 0
    tr = TValueTrainer(succ, fail, seed=1)
    x, y = tr._sample()
    z = tr.net.forward(x, train=True).clone()
    tr.stats.zero_()
    from seqdex_b200 import _lib
    from seqdex_b200.ppo import _p, _stream
    _lib.check(tr.L.sdx_tvalue_bce(_p(z), _p(y), tr.batch, _p(tr.dz), _p(tr.stats), _stream()))
    zt = z.clone().requires_grad_(True)
    tgt = torch.nn.functional.one_hot(y.long(), 2).float()
    loss = torch.nn.BCEWithLogitsLoss()(torch.nn.functional.elu(zt), tgt)
    loss.backward()
    torch.cuda.synchronize()
    torch.testing.assert_close(tr.stats[0] / (2 * tr.batch), loss.detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(tr.dz, zt.grad, rtol=1e-4, atol=1e-8)
    acc0 = tr.validate()
    tr.train_rollout(300)
    acc = tr.validate()
    assert acc > 0.9 and acc >= acc0, (acc0, acc)
    from seqdex_b200.env import SdxEnv
    e = SdxEnv(scene, 8)
    e.set_tvalue_weights(tr.weights())


def test_ppo_learns_on_the_cuda_env(scene):
    """the whole loop -- contact step, observations, reward, PPO update on the tensor-core MLPs -- learns: the mean rollout reward of
    BlockAssemblyGraspSim rises by more than half within 200 iterations at 2048 envs (typically x 5 - x 10) (profiles/r01_learning_curve_*.txt hold
    the long curves of all three tasks: 0.005 -> 7.8 in 600 iterations for GraspSim)"""
    from seqdex_b200.ppo import A2CAgent, PPOConfig
    from seqdex_b200.tasks import BlockAssemblyGraspSim
    from seqdex_b200.vec_task import RLgamesVecTaskPython
    cfg = {"env": {"numEnvs": 2048, "episodeLength": 150, "actionsMovingAverage": 1.0}, "sim": {"substeps": 2, "physx": {}}, "task": {"randomize": False}}
    task = BlockAssemblyGraspSim(cfg, bank_per_type=8)
    agent = A2CAgent(RLgamesVecTaskPython(task, "cuda:0"), PPOConfig(minibatch_size=4096))
    rew = [agent.train_epoch()["mean_reward"] for _ in range(200)]
    first, last = sum(rew[:15]) / 15, sum(rew[-15:]) / 15
    assert all(math.isfinite(r) for r in rew)
    # The update is not bit-reproducible run to run (split-K float atomics), so the curve's early slope varies: over ten
    # runs of the earlier 150-iteration version the ratio last / first ranged from 2.6 to 9.  The bound (x 1.6 after 200 iterations) is what every run clears with margin; an agent that does not learn stays within 10 % of `first`.
    assert last > 1.6 * first and last > 0.011, (first, last)


def test_pipelined_backward_publishes_the_same_gradients_layer_by_layer():
    """sdx_mlp_backward_pipelined (per-layer unpack + event, what the data-parallel agent all-reduces layer by layer) == sdx_mlp_backward;
    a side stream that waits for layer l's event sees that layer's final slice while the layers below are still being differentiated"""
    from seqdex_b200.ppo import MLP
    torch.manual_seed(6)
    m = MLP(396, 23, 2048, has_sigma=True, seed=4)
    x = torch.randn(2048, 396, device="cuda")
    dout = torch.randn(2048, 23, device="cuda") / 2048
    m.forward(x, train=True)
    m.backward(dout)
    ref = m.grads.clone()
    m.grads.zero_()
    m.forward(x, train=True)
    side = torch.cuda.Stream()
    snap = {}
    m.backward_pipelined(dout)
    for layer in (3, 2, 1, 0):
        b, e = m.layer_range(layer)
        m.wait_layer(layer, side)
        with torch.cuda.stream(side):
            snap[layer] = (b, e, m.grads[b:e].clone())
    torch.cuda.synchronize()
    n = m.nparams - 23                       # sigma is the loss kernel's business
    torch.testing.assert_close(m.grads[:n], ref[:n], rtol=1e-4, atol=1e-6)          # split-K atomics: order-dependent last bits
    ends = []
    for layer, (b, e, g) in snap.items():
        torch.testing.assert_close(g, ref[b:e], rtol=1e-4, atol=1e-6)
        ends.append((b, e))
    assert sorted(ends)[0][0] == 0 and sorted(ends)[-1][1] == n and all(sorted(ends)[i][1] == sorted(ends)[i + 1][0] for i in range(3))


def test_update_as_cuda_graph_follows_the_eager_update(monkeypatch):
    """A2CAgent.update captures its ~800 launches into one CUDA graph from the second call on (the learning rate, the Adam step and the
    statistics live on the device).  The dW GEMMs reduce their split-K partials with float atomics, so two runs of the SAME path already
    differ in the last bits of the gradients (and Adam turns a sign flip of a near-zero gradient into a full step): the check is that the
    graph path counts its optimiser steps on the device -- a graph with host-side counters would freeze the bias corrections --,
    keeps the adaptive learning rate alive, and moves the parameters the way the eager path does."""
    from seqdex_b200.ppo import A2CAgent, PPOConfig
    from seqdex_b200.tasks import BlockAssemblyGraspSim
    from seqdex_b200.vec_task import RLgamesVecTaskPython
    cfg = {"env": {"numEnvs": 256, "episodeLength": 150, "actionsMovingAverage": 1.0}, "sim": {"substeps": 2, "physx": {}}, "task": {"randomize": False}}
    out = {}
    for mode in ("0", "force", "0b"):
        monkeypatch.setenv("SEQDEX_PPO_GRAPH", mode[0] if mode != "force" else mode)
        torch.manual_seed(17)                       # VecTask.reset draws its first actions from the global generator
        task = BlockAssemblyGraspSim(cfg, bank_per_type=2)
        agent = A2CAgent(RLgamesVecTaskPython(task, "cuda:0"), PPOConfig(minibatch_size=512))
        w0 = (agent.actor.params.clone(), agent.cv.params.clone())
        infos = [agent.train_epoch() for _ in range(6)]
        torch.cuda.synchronize()
        assert (getattr(agent, "_graph", None) is not None) == (mode == "force")
        out[mode] = (agent.actor.params - w0[0], agent.cv.params - w0[1], infos, agent._adam_step(agent.actor), agent._adam_step(agent.cv))
        task.env.close()
    cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm()))
    a, b, a2 = out["0"], out["force"], out["0b"]
    assert a[3] == b[3] == 6 * 5 * 4 and a[4] == b[4] == 6 * 5 * 4        # 6 iterations x 5 mini-epochs x 4 minibatches, counted on the device
    assert all(math.isfinite(i["kl"]) and 1e-6 <= i["lr"] <= 1e-2 for i in b[2])
    assert len({i["lr"] for i in b[2]}) > 1                               # the device-side schedule keeps adapting under replay
    # graph vs eager agree as well as eager vs eager does (the run-to-run spread of the atomics), within a factor
    ref_a, ref_c = cos(a[0], a2[0]), cos(a[1], a2[1])
    assert cos(a[0], b[0]) > 0.8 * ref_a and cos(a[1], b[1]) > 0.8 * ref_c, (cos(a[0], b[0]), ref_a, cos(a[1], b[1]), ref_c)
    assert 0.5 < float(b[0].norm() / a[0].norm()) < 2.0
