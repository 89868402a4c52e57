"""PPO with rl_games 1.5.2 semantics (continuous_a2c_logstd + asymmetric central value;
cfg/lego/ppo_continuous_grasp.yaml) on top of the CUDA kernels of csrc/sdx_ppo.cu:
every dense contraction is the tcgen05 GEMM, every elementwise PPO op a fused kernel.  This module only
sequences launches (the role rl_games' A2CAgent.play_steps / train_epoch plays; restated in-tree at
utils/rl_games_custom.py RGC:1394-1483, 1621-1683, 1767-1911) and owns no math of its own.

Multi-GPU: envs are sharded over ranks; the ONLY data-path collectives are one all-reduce of the flat fp32
gradient vector per optimiser step, one of the advantage moments per iteration, one of the RunningMeanStd
column moments per iteration and one scalar (KL) per mini-epoch, all through torch.distributed (NCCL).
"""
from __future__ import annotations

import contextlib
import ctypes
import math
import os

import torch

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


GRAD_TAIL = 16        # SDX_GRAD_TAIL (include/seqdex_b200.h)
LR_MIN, LR_MAX = 1e-6, 1e-2   # rl_games AdaptiveScheduler bounds


def adaptive_lr(lr, kl, kl_threshold):
    """rl_games 1.5.2 ``AdaptiveScheduler.update`` (called at RGC:1360-1365 after every minibatch: the SeqDex yamls leave
    ``schedule_type`` at its default 'legacy').  Host restatement of what ``sdx_ppo_adaptive_lr`` does on the device; used where a
    learning rate lives on the host (checkpoint restore, tests/test_ppo_oracle_golden.py)."""
    if kl > 2.0 * kl_threshold:
        lr = max(lr / 1.5, LR_MIN)
    if kl < 0.5 * kl_threshold:
        lr = min(lr * 1.5, LR_MAX)
    return lr


class _View:
    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class MLP:
    """in -> 1024 -> 512 -> 256 -> out, ELU (cfg/lego/ppo_continuous_grasp.yaml:21-23): fp32 master params,
    bf16 tensor-core compute.  ``params`` / ``grads`` are flat fp32 views in torch state_dict order."""

    def __init__(self, in_dim, out_dim, max_rows, has_sigma=False, device=0, seed=0, hidden=(1024, 512, 256)):
        self.L = _lib.load()
        self.in_dim, self.out_dim, self.max_rows, self.has_sigma = in_dim, out_dim, max_rows, has_sigma
        self.in_pad = (in_dim + 63) // 64 * 64
        self.device = torch.device("cuda", device)
        self.h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.L.sdx_mlp_create_ex(in_dim, out_dim, hidden[0], hidden[1], hidden[2], max_rows, int(has_sigma), ctypes.byref(self.h)))
            n = ctypes.c_int64()
            pp, pg, po, pm, pv = (ctypes.c_void_p() for _ in range(5))
            _lib.check(self.L.sdx_mlp_info(self.h, ctypes.byref(n), ctypes.byref(pp), ctypes.byref(pg), ctypes.byref(po),
                                           ctypes.byref(pm), ctypes.byref(pv)))
            self.nparams = n.value
            mk = lambda p, shp: torch.as_tensor(_View(p.value, shp), device=self.device)
            self.params, self.grads = mk(pp, [self.nparams]), mk(pg, [self.nparams])
            # the gradient buffer carries GRAD_TAIL extra floats: loss statistics written there are summed by the SAME all-reduce
            self.grads_ext = mk(pg, [self.nparams + GRAD_TAIL])
            self.stats = self.grads_ext[self.nparams:self.nparams + 4]
            self.adam_m, self.adam_v = mk(pm, [self.nparams]), mk(pv, [self.nparams])
            self.out = mk(po, [max_rows, out_dim])
        self.dims = [in_dim, hidden[0], hidden[1], hidden[2], out_dim]
        self.init_default(seed)

    def close(self):
        if self.h:
            self.L.sdx_mlp_destroy(self.h)
            self.h = ctypes.c_void_p()

    def slices(self):
        """[(name, offset, shape)] of the flat parameter vector (torch state_dict order)"""
        out, off = [], 0
        for l in range(4):
            o, i = self.dims[l + 1], self.dims[l]
            out.append((f"W{l}", off, (o, i))); off += o * i
            out.append((f"b{l}", off, (o,))); off += o
        if self.has_sigma:
            out.append(("sigma", off, (self.out_dim,))); off += self.out_dim
        assert off == self.nparams
        return out

    def init_default(self, seed=0):
        """rl_games 'default' initialiser = torch.nn.Linear's (U(-1/sqrt(in), 1/sqrt(in)) for W and b); sigma const 0"""
        g = torch.Generator().manual_seed(seed)
        flat = torch.zeros(self.nparams)
        for name, off, shp in self.slices():
            if name == "sigma":
                continue
            fan_in = self.dims[int(name[1])]
            n = int(torch.tensor(shp).prod())
            flat[off:off + n] = (torch.rand(n, generator=g) * 2 - 1) / math.sqrt(fan_in)
        self.load_flat(flat)

    def load_flat(self, flat):
        self.params.copy_(flat.to(self.device, torch.float32))
        self.sync()

    def sync(self):
        _lib.check(self.L.sdx_mlp_sync(self.h, _stream()))

    def forward(self, x, mean=None, var=None, train=False):
        M = x.shape[0]
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == self.in_dim
        _lib.check(self.L.sdx_mlp_forward(self.h, _p(x), M, _p(mean), _p(var), int(train), _stream()))
        return self.out[:M]

    def convert_batch(self, x, xb, xt, mean=None, var=None):
        """fp32 [B, in] -> bf16 [B, in_pad] + transposed bf16 [in_pad + 16, B] (ones row appended), once per iteration"""
        _lib.check(self.L.sdx_mlp_convert_batch(self.h, _p(x), x.shape[0], _p(mean), _p(var), _p(xb), _p(xt), _stream()))

    def convert_batch_env_major(self, x, horizon, xb, xt, mean=None, var=None):
        """x: TIME-major rollout buffer [H * N, in]; rows of the converted batch are ENV-major (rl_games' swap_and_flatten01)"""
        _lib.check(self.L.sdx_mlp_convert_batch_env_major(self.h, _p(x), x.shape[0], int(horizon), _p(mean), _p(var), _p(xb), _p(xt), _stream()))

    def forward_pre(self, xb, xt, row0, M, train=True):
        _lib.check(self.L.sdx_mlp_forward_pre(self.h, _p(xb), _p(xt), xb.shape[0], row0, M, int(train), _stream()))
        return self.out[:M]

    def backward(self, dout):
        assert dout.is_contiguous() and dout.dtype == torch.float32
        _lib.check(self.L.sdx_mlp_backward(self.h, _p(dout), dout.shape[0], _stream()))

    def backward_pipelined(self, dout):
        """same gradients, published layer by layer (output layer first) with an event per layer: see ``wait_layer``"""
        assert dout.is_contiguous() and dout.dtype == torch.float32
        _lib.check(self.L.sdx_mlp_backward_pipelined(self.h, _p(dout), dout.shape[0], _stream()))

    def wait_layer(self, layer, stream):
        """make ``stream`` wait until layer ``layer``'s slice of ``grads`` is final (after the last backward_pipelined)"""
        _lib.check(self.L.sdx_mlp_wait_layer(self.h, int(layer), ctypes.c_void_p(stream.cuda_stream)))

    def layer_range(self, layer):
        b, e = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.L.sdx_mlp_layer_range(self.h, int(layer), ctypes.byref(b), ctypes.byref(e)))
        return int(b.value), int(e.value)

    def adam(self, lr, max_norm=1.0, b1=0.9, b2=0.999, eps=1e-8):
        _lib.check(self.L.sdx_mlp_adam(self.h, ctypes.c_float(lr), ctypes.c_float(b1), ctypes.c_float(b2), ctypes.c_float(eps),
                                       ctypes.c_float(max_norm), _stream()))

    def adam_dev(self, lr_dev, max_norm=1.0, b1=0.9, b2=0.999, eps=1e-8):
        """Adam step whose learning rate is read from device memory when the kernel runs"""
        _lib.check(self.L.sdx_mlp_adam_dev(self.h, _p(lr_dev), ctypes.c_float(b1), ctypes.c_float(b2), ctypes.c_float(eps),
                                           ctypes.c_float(max_norm), _stream()))

    def torch_reference(self):
        """plain fp32 torch modules holding the same parameters (tests / checkpoint export)"""
        layers = []
        sl = {n: (o, s) for n, o, s in self.slices()}
        for l in range(4):
            lin = torch.nn.Linear(self.dims[l], self.dims[l + 1]).to(self.device)
            o, s = sl[f"W{l}"]
            lin.weight.data.copy_(self.params[o:o + s[0] * s[1]].view(s))
            o, s = sl[f"b{l}"]
            lin.bias.data.copy_(self.params[o:o + s[0]])
            layers.append(lin)
            if l < 3:
                layers.append(torch.nn.ELU())
        return torch.nn.Sequential(*layers)


class PPOConfig:
    """cfg/lego/ppo_continuous_grasp.yaml:30-95 (minibatch_size defaults to a sane value instead of the yaml's 4,
    SURVEY.md section 7 'hard parts'; pass minibatch_size=4 to honour the yaml literally)"""

    def __init__(self, **kw):
        self.gamma, self.tau = 0.99, 0.95
        self.learning_rate, self.cv_learning_rate = 3e-4, 1e-3
        self.grad_norm, self.e_clip, self.clip_value = 1.0, 0.1, True
        self.horizon_length, self.mini_epochs, self.cv_mini_epochs = 8, 5, 5
        self.minibatch_size = 16384
        self.kl_threshold, self.bounds_loss_coef = 0.02, 0.001
        self.normalize_advantage, self.cv_normalize_input = True, True
        self.lr_schedule = "adaptive"
        # data parallel: all-reduce each layer's gradients as soon as its dW GEMM retires (overlapping the layers below) instead of one
        # exchange after the whole backward.  MEASURED SLOWER at these sizes (2 x B200: 7.38 vs 7.22 ms/step -- four latency-bound NCCL
        # launches of <= 2 MB instead of one of 8.5 MB), so it is off by default (profiles/r02_allreduce_pipelining.txt)
        self.pipeline_allreduce = False
        self.seed = 22
        for k, v in kw.items():
            if not hasattr(self, k):
                raise KeyError(k)
            setattr(self, k, v)


class A2CAgent:
    """the calls rl_games' Runner makes on its agent: play_steps() + train_epoch() (= train() loop body)."""

    def __init__(self, vec_env, cfg: PPOConfig | None = None, device=0, dist_group=None):
        self.env, self.cfg = vec_env, cfg or PPOConfig()
        c = self.cfg
        self.device = torch.device("cuda", device)
        self.N = vec_env.num_envs
        self.H = c.horizon_length
        self.B = self.N * self.H
        self.mb = min(c.minibatch_size, self.B)
        assert self.B % self.mb == 0, "batch must be a multiple of the minibatch"
        self.A, self.obs_dim, self.state_dim = vec_env.num_actions, vec_env.num_obs, vec_env.num_states
        rows = max(self.N, self.mb)
        self.L = _lib.load()
        self.actor = MLP(self.obs_dim, self.A, rows, has_sigma=True, device=device, seed=c.seed)
        self.cv = MLP(self.state_dim, 1, rows, has_sigma=False, device=device, seed=c.seed + 1)
        self.dist = dist_group
        self.world = torch.distributed.get_world_size(dist_group) if dist_group is not None else 1
        self.rank = torch.distributed.get_rank(dist_group) if dist_group is not None else 0
        # parameters start identical on every rank (same init seed); the exploration noise must NOT: rank r's env e would
        # otherwise draw exactly rank 0's env e noise (Philox key = env index)
        self.sample_seed = c.seed + 0x9E3779B1 * self.rank
        z = lambda *s, dt=torch.float32: torch.zeros(*s, device=self.device, dtype=dt)
        H, N, A = self.H, self.N, self.A
        self.b_obs, self.b_states = z(H, N, self.obs_dim), z(H, N, self.state_dim)
        self.b_actions, self.b_mu = z(H, N, A), z(H, N, A)
        self.b_neglogp, self.b_values, self.b_rewards, self.b_dones = z(H, N), z(H, N), z(H, N), z(H, N)
        self.b_adv, self.b_returns = z(H, N), z(H, N)
        self.old_logstd = z(self.B // self.mb, A)
        bf = lambda r, c: torch.zeros(r, c, device=self.device, dtype=torch.bfloat16)
        self.xb_obs, self.xt_obs = bf(self.B, self.actor.in_pad), bf(self.actor.in_pad + 16, self.B)
        self.xb_st, self.xt_st = bf(self.B, self.cv.in_pad), bf(self.cv.in_pad + 16, self.B)
        self.dmu, self.dv = z(self.mb, A), z(self.mb, 1)
        self.stats, self.cv_stats = self.actor.stats, z(4)      # actor statistics live in the tail of the gradient buffer
        self.lr_dev = torch.full((1,), float(c.learning_rate), device=self.device)
        self.accum = z(8)                                      # per-iteration sums of the minibatch statistics | #minibatches | last kl
        self.mom = torch.zeros(2, device=self.device, dtype=torch.float64)
        self.colmom = torch.zeros(2 * self.state_dim, device=self.device, dtype=torch.float64)
        self.rms_mean, self.rms_var = z(self.state_dim), torch.ones(self.state_dim, device=self.device)
        self.rms_count = torch.full((1,), 1e-4, device=self.device, dtype=torch.float64)
        self.last_lr = c.learning_rate
        self.dones = torch.zeros(N, device=self.device)
        self.last_values = torch.zeros(N, device=self.device)
        self.obs = None
        self.sample_counter = 0
        self.epoch_num = 0
        self.last_kl = 0.0

    def _side_stream(self):
        """second stream for the central-value update (SEQDEX_PPO_STREAMS=0 keeps everything on one stream)"""
        if os.environ.get("SEQDEX_PPO_STREAMS", "1") == "0":
            return None
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    @property
    def logstd(self):
        return self.actor.params[self.actor.nparams - self.A:]

    # ---- rollout (RGC:1394-1483)
    def get_action_values(self, obs, states):
        mu = self.actor.forward(obs)
        t = self.sample_counter
        self.sample_counter += 1
        return mu, t

    # the three pieces of one rollout step, so that a sequencing runner (PSR:220-275) can interleave two agents on one env
    def act(self, t):
        """store obs / states / dones of step t (RGC:1403-1410, PSR:338-346), sample the action, evaluate the central value"""
        L, A, N = self.L, self.A, self.N
        mean = self.rms_mean if self.cfg.cv_normalize_input else None
        var = self.rms_var if self.cfg.cv_normalize_input else None
        self.b_obs[t].copy_(self.next_obs)
        self.b_states[t].copy_(self.next_states)
        self.b_dones[t].copy_(self.dones)
        mu = self.actor.forward(self.b_obs[t])
        self.b_mu[t].copy_(mu)
        _lib.check(L.sdx_ppo_sample(_p(self.b_mu[t]), _p(self.logstd), N, A, ctypes.c_uint64(self.sample_seed), self.sample_counter,
                                    _p(self.b_actions[t]), _p(self.b_neglogp[t]), _stream()))
        self.sample_counter += 1
        v = self.cv.forward(self.b_states[t], mean, var)
        self.b_values[t].copy_(v.view(-1))
        return self.b_actions[t]

    def record(self, t, rew, dones):
        """rewards / dones after the env step (RGC:1420-1440, PSR:348-354; reward_shaper scale 1, no time-out bootstrap)"""
        self.b_rewards[t].copy_(rew)
        self.dones.copy_(dones)

    def finish_rollout(self):
        """bootstrap value + GAE (RGC:1465-1478, PSR:322-336)"""
        mean = self.rms_mean if self.cfg.cv_normalize_input else None
        var = self.rms_var if self.cfg.cv_normalize_input else None
        v = self.cv.forward(self.next_states, mean, var)
        self.last_values.copy_(v.view(-1))
        _lib.check(self.L.sdx_gae(_p(self.b_rewards), _p(self.b_values), _p(self.b_dones), _p(self.last_values), _p(self.dones), _p(self.b_adv),
                                  _p(self.b_returns), self.H, self.N, ctypes.c_float(self.cfg.gamma), ctypes.c_float(self.cfg.tau), _stream()))

    def set_obs(self, obs, states):
        """the observation the next act() consumes (``agent.obs`` in rl_games)"""
        if self.obs is None:
            self.next_obs, self.next_states = obs.clone(), states.clone()
            self.obs = True
        else:
            self.next_obs.copy_(obs); self.next_states.copy_(states)

    def play_steps(self):
        """horizon_length env steps (RGC:1394-1483): obs/dones stored pre-step, values from the central value net"""
        fast = hasattr(self.env, "step_into")
        if self.obs is None:
            first = self.env.reset()
            self.set_obs(first["obs"], first["states"])
        for t in range(self.H):
            a = self.act(t)
            if fast:
                rew, dones, _ = self.env.step_into(a, self.next_obs, self.next_states)
            else:
                o, rew, dones, _ = self.env.step(a)
                self.next_obs.copy_(o["obs"]); self.next_states.copy_(o["states"])
            self.record(t, rew, dones)
        self.finish_rollout()

    def _allreduce(self, t, avg=True):
        if self.dist is not None and self.world > 1:
            from .dist_utils import allreduce_
            allreduce_(t, self.dist, avg)

    def _backward_and_allreduce(self, mlp, dout, tail=0):
        """backward + gradient exchange.  One rank: plain backward.  Several: the backward publishes its gradients layer by layer
        (output layer first) and each layer's slice is all-reduced -- on a communication stream that waits only for THAT layer's
        event -- while the layers below are still being differentiated; the output layer's bucket carries everything behind it in
        the flat vector (sigma and ``tail`` statistics floats, written by the loss kernel before the backward started).  The
        compute stream waits for the communication stream before the optimiser step."""
        if self.dist is None or self.world == 1:
            mlp.backward(dout)
            return
        if not self.cfg.pipeline_allreduce:
            mlp.backward(dout)
            self._allreduce(mlp.grads_ext[:mlp.nparams + tail])
            return
        cur = torch.cuda.current_stream()
        comm = self._comm_stream(cur)
        mlp.backward_pipelined(dout)
        end_all = mlp.nparams + tail
        for layer in (3, 2, 1, 0):
            b, e = mlp.layer_range(layer)
            if layer == 3:
                e = end_all       # sigma gradient / statistics: written by the loss kernel BEFORE the backward on `cur`, so layer 3's event covers them
            mlp.wait_layer(layer, comm)
            with torch.cuda.stream(comm):
                self._allreduce(mlp.grads_ext[b:e])
        cur.wait_stream(comm)

    def _comm_stream(self, cur):
        """one communication stream per compute stream (the actor and the central-value chains run on two streams)"""
        d = self.__dict__.setdefault("_comm_streams", {})
        k = cur.cuda_stream
        if k not in d:
            d[k] = torch.cuda.Stream(device=self.device)
        return d[k]

    # ---- update (RGC:1621-1683, 1339-1375, 1767-1911)
    def train_epoch(self):
        self.play_steps()
        return self.update()

    def update(self):
        """prepare_dataset + central-value and actor mini-epochs on the rollout the buffers hold (RGC:1621-1683, 1306-1392, PSR:277-319).
        The batch is ENV-major as in rl_games (``swap_and_flatten01``, RGC:1480-1481): a minibatch is a contiguous block of envs with
        all H steps of each.  The adaptive-KL schedule runs after EVERY minibatch (``schedule_type`` defaults to 'legacy',
        RGC:1360-1365), on the device.

        The update is some 800 kernel launches issued from Python.  Nothing in it depends on the host (the learning rate, the Adam step
        and the statistics live on the device), so on one GPU it is CAPTURED ONCE into a CUDA graph from the second call on and
        replayed: at small env counts, where the update is launch-bound, that is 6.2 -> 3.9 ms per update (256 envs); at 16 384 envs the
        GEMMs dominate and it is 15.4 -> 14.8 ms (tools/ppo_graph_check.py).  ``SEQDEX_PPO_GRAPH=0`` keeps the eager path (identical
        results, tests/test_ppo_gpu.py); with several ranks the update stays eager unless ``SEQDEX_PPO_GRAPH=force`` (its all-reduces
        would have to be captured by the communicator)."""
        self.lr_dev.fill_(self.last_lr)
        mode = os.environ.get("SEQDEX_PPO_GRAPH", "1")
        graph_ok = (mode != "0" and not getattr(self, "_graph_failed", False)
                    and (self.world == 1 or (mode == "force" and not self.cfg.pipeline_allreduce)))
        if graph_ok and getattr(self, "_graph", None) is not None:
            self._graph.replay()
            self.L.sdx_ppo_add_launches(ctypes.c_longlong(self._graph_launches))
        elif graph_ok and getattr(self, "_eager_updates", 0) >= 1:
            self.L.sdx_ppo_launch_count.restype = ctypes.c_longlong
            torch.cuda.synchronize()
            l0 = int(self.L.sdx_ppo_launch_count())
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g, capture_error_mode="thread_local"):   # other threads (NCCL's watchdog, a clock sampler) keep running
                    self._update_body()
            except Exception:
                if mode == "force":
                    raise
                # capture is an optimisation: this agent falls back to the eager path for good (e.g. a collective the communicator cannot capture)
                torch.cuda.synchronize()
                self._graph_failed = True
                self._update_body()
            else:
                self._graph, self._graph_launches = g, int(self.L.sdx_ppo_launch_count()) - l0
                g.replay()                                     # capturing does not execute
        else:
            self._update_body()
            self._eager_updates = getattr(self, "_eager_updates", 0) + 1
        acc = self.accum.tolist() + [float(self.lr_dev)]          # the iteration's one host read
        nb = max(acc[4], 1.0) * self.mb
        self.last_kl, self.last_lr = acc[5], acc[8]
        self.epoch_num += 1
        return {"kl": self.last_kl, "lr": self.last_lr, "a_loss": acc[0] / nb, "b_loss": acc[1] / nb, "kl_mean": acc[2] / nb,
                "mean_reward": float(self.b_rewards.mean())}

    def _update_body(self):
        """everything of update() that runs on the device, free of host reads (capturable)"""
        c, L, A, B, mb, H = self.cfg, self.L, self.A, self.B, self.mb, self.H
        Btot = B * self.world
        em = lambda t: t.transpose(0, 1).reshape(B, *t.shape[2:]).contiguous()          # swap_and_flatten01
        actions, mu_old, nlp_old = em(self.b_actions), em(self.b_mu), em(self.b_neglogp)
        values, returns = em(self.b_values), em(self.b_returns)
        adv = (returns - values).contiguous()
        states_tm = self.b_states.view(B, -1)
        if c.normalize_advantage:                       # (adv - mean) / (std + 1e-8) over the GLOBAL batch (RGC:1651)
            _lib.check(L.sdx_moments(_p(adv), B, _p(self.mom), _stream()))
            self._allreduce(self.mom, avg=False)
            _lib.check(L.sdx_normalize(_p(adv), B, _p(self.mom), ctypes.c_double(Btot), _stream()))
        if c.cv_normalize_input:                        # RunningMeanStd of the critic state, merged once per iteration
            _lib.check(L.sdx_col_moments(_p(states_tm), B, self.state_dim, _p(self.colmom), _stream()))
            self._allreduce(self.colmom, avg=False)
            _lib.check(L.sdx_rms_merge(_p(self.rms_mean), _p(self.rms_var), _p(self.rms_count), _p(self.colmom), self.state_dim,
                                       ctypes.c_double(Btot), _stream()))
        nmb = B // mb
        inv = 1.0 / float(mb)
        # inputs of both networks -> bf16 (row-major + transposed, env-major rows) ONCE per iteration; minibatches are slices of these
        self.actor.convert_batch_env_major(self.b_obs.view(B, -1), H, self.xb_obs, self.xt_obs)
        self.cv.convert_batch_env_major(states_tm, H, self.xb_st, self.xt_st, self.rms_mean if c.cv_normalize_input else None,
                                        self.rms_var if c.cv_normalize_input else None)
        # The two networks' updates are independent once the rollout is in the buffers: the central-value chain goes to a side
        # stream so that its small kernels (loss, norm, Adam, unpack) and GEMM tails overlap the actor chain's GEMMs and vice
        # versa.  The steps of the two chains are ISSUED alternately (so that, with several GPUs, their all-reduces enter the
        # communicator interleaved instead of one chain queueing behind the other); each chain runs in program order on its
        # own stream and the results do not depend on the interleaving.
        side = self._side_stream()
        main = torch.cuda.current_stream()
        if side is not None:
            side.wait_stream(main)
        side_ctx = (lambda: torch.cuda.stream(side)) if side is not None else contextlib.nullcontext
        na = self.actor.nparams

        def cv_step(i):       # central value network (asymmetric critic), own optimiser lr 1e-3
            s = slice(i * mb, (i + 1) * mb)
            v = self.cv.forward_pre(self.xb_st, self.xt_st, i * mb, mb)
            _lib.check(L.sdx_ppo_value_loss(_p(v), _p(values[s]), _p(returns[s]), mb, ctypes.c_float(c.e_clip), int(c.clip_value),
                                            ctypes.c_float(inv), _p(self.dv), _p(self.cv_stats), _stream()))
            self._backward_and_allreduce(self.cv, self.dv)
            self.cv.adam(c.cv_learning_rate, c.grad_norm)

        def actor_step(i):
            s = slice(i * mb, (i + 1) * mb)
            mu = self.actor.forward_pre(self.xb_obs, self.xt_obs, i * mb, mb)
            self.actor.grads[na - A:].zero_()
            _lib.check(L.sdx_ppo_actor_loss(_p(mu), _p(self.logstd), _p(actions[s]), _p(mu_old[s]), _p(self.old_logstd[i]), _p(nlp_old[s]),
                                            _p(adv[s]), mb, A, ctypes.c_float(c.e_clip), ctypes.c_float(c.bounds_loss_coef),
                                            ctypes.c_float(inv), _p(self.dmu), _p(self.actor.grads[na - A:]),
                                            _p(self.stats), _stream()))
            mu_old[s].copy_(mu)                                  # dataset.update_mu_sigma (RGC:1358)
            self.old_logstd[i].copy_(self.logstd)
            self._backward_and_allreduce(self.actor, self.dmu, tail=4)   # gradients AND the minibatch statistics (KL averaged over ranks, RGC:1361-1362)
            self.actor.adam_dev(self.lr_dev, c.grad_norm)
            _lib.check(L.sdx_ppo_adaptive_lr(_p(self.stats), ctypes.c_float(inv), ctypes.c_float(c.kl_threshold), ctypes.c_float(LR_MIN),
                                             ctypes.c_float(LR_MAX), _p(self.lr_dev), _p(self.accum), int(c.lr_schedule == "adaptive"), _stream()))

        self.old_logstd.copy_(self.logstd.unsqueeze(0).expand(nmb, A))
        self.accum.zero_()
        self.stats.zero_()
        for ep in range(max(c.mini_epochs, c.cv_mini_epochs)):
            for i in range(nmb):
                if ep < c.cv_mini_epochs:
                    with side_ctx():
                        cv_step(i)
                if ep < c.mini_epochs:
                    actor_step(i)
        if side is not None:
            main.wait_stream(side)

    # ---- checkpoint (rl_games .pth layout, seqdex_b200/checkpoint.py; RGC:1913-1933, 2098-2106)
    def _adam_step(self, mlp, set_to=-1):
        self.L.sdx_mlp_adam_step.restype = ctypes.c_longlong
        return int(self.L.sdx_mlp_adam_step(mlp.h, ctypes.c_longlong(set_to)))

    def get_weights(self):
        from . import checkpoint as ck
        return {"model": ck.actor_state_dict(self.actor.params, self.obs_dim, self.A, critic=getattr(self, "_a2c_critic", None), seed=self.cfg.seed)}

    def get_full_state_weights(self):
        """what rl_games' ``A2CBase.save`` writes (restated at RGC:1913-1933): weights + optimiser + central value + counters"""
        from . import checkpoint as ck
        state = self.get_weights()
        state["epoch"] = self.epoch_num
        ent = ck.a2c_param_entries(self.obs_dim, self.A)
        state["optimizer"] = ck.adam_state_dict(self.actor.adam_m, self.actor.adam_v, self._adam_step(self.actor), self.actor.slices(),
                                                self.last_lr, [n for n, _ in ent], shapes=dict(ent))
        state["assymetric_vf_nets"] = ck.central_value_state_dict(
            self.cv.params, self.state_dim, rms=(self.rms_mean, self.rms_var, self.rms_count) if self.cfg.cv_normalize_input else None)
        cvo = [n for n, _, _ in self.cv.slices()]
        state["assymetric_vf_optimizer"] = ck.adam_state_dict(self.cv.adam_m, self.cv.adam_v, self._adam_step(self.cv), self.cv.slices(),
                                                              self.cfg.cv_learning_rate, cvo)
        state["frame"] = self.epoch_num * self.B * self.world
        state["last_mean_rewards"] = getattr(self, "last_mean_rewards", -100500)
        state["env_state"] = self.env.get_env_state() if hasattr(self.env, "get_env_state") else None
        return state

    def set_weights(self, weights):
        from . import checkpoint as ck
        flat, critic = ck.actor_flat(weights["model"], self.obs_dim, self.A)
        self._a2c_critic = critic
        self.actor.load_flat(flat)

    def set_full_state_weights(self, weights, load_optimizer_state=True):
        from . import checkpoint as ck
        self.set_weights(weights)
        self.epoch_num = int(weights.get("epoch", 0))
        self.last_mean_rewards = weights.get("last_mean_rewards", -100500)
        if "assymetric_vf_nets" in weights:
            flat, rms = ck.central_value_flat(weights["assymetric_vf_nets"], self.state_dim)
            self.cv.load_flat(flat)
            if rms is not None:
                self.rms_mean.copy_(rms[0]); self.rms_var.copy_(rms[1]); self.rms_count.copy_(rms[2])
        if load_optimizer_state:
            for mlp, key, order in ((self.actor, "optimizer", ck.actor_param_order()), (self.cv, "assymetric_vf_optimizer", [n for n, _, _ in self.cv.slices()])):
                opt = weights.get(key)
                if opt and key == "optimizer" and len(opt["state"]) == len(ck.LEGACY_ACTOR_ORDER):
                    order = ck.LEGACY_ACTOR_ORDER
                if not opt or len(opt["state"]) != len(order):
                    continue
                by = {n: (o, s) for n, o, s in mlp.slices()}
                for i, name in enumerate(order):
                    if name not in by:
                        continue    # the a2c net's own critic trunk / value head: rl_games' optimiser holds them, this engine does not train them
                    o, shp = by[name]
                    k = opt["state"][i]["exp_avg"].numel()
                    mlp.adam_m[o:o + k].copy_(opt["state"][i]["exp_avg"].reshape(-1))
                    mlp.adam_v[o:o + k].copy_(opt["state"][i]["exp_avg_sq"].reshape(-1))
                self._adam_step(mlp, int(float(opt["state"][0]["step"])))
                if key == "optimizer":
                    self.last_lr = float(opt["param_groups"][0]["lr"])

    def save(self, fn):
        from . import checkpoint as ck
        return ck.save_checkpoint(fn, self.get_full_state_weights())

    def restore(self, fn):
        from . import checkpoint as ck
        self.set_full_state_weights(ck.load_checkpoint(fn))

    def state_dict(self):
        return self.get_full_state_weights()
