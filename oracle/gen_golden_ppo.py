#!/usr/bin/env python
"""Golden vectors for the rl_games half of the hot path by EXECUTING THE REFERENCE'S OWN
PYTHON (build container only: it reads /root/reference; the outputs are committed under tests/golden/).

``utils/rl_games_custom.py`` (RGC) cannot be imported -- its module body needs rl_games, and names rl_games would have brought
in -- so the functions below are compiled from the reference's source text (``ast`` nodes of the file where it lies, decorators
dropped) and executed unmodified:

    _calc_neglogp                    RGC:2113-2127   -> ppo_neglogp.npz
    _calc_ac_loss                    RGC:2129-2132   -> ppo_ac_loss.npz
    A2CControllerAgent.play_steps    RGC:1394-1483   -> ppo_play_steps.npz   (what is stored when, what reaches discount_values)
    A2CControllerAgent.prepare_dataset RGC:1621-1683 -> ppo_prepare_dataset.npz (advantages = returns - values, normalisation)
    A2CControllerAgent.train_epoch   RGC:1306-1392   -> ppo_schedule.npz     (order of minibatches, update_mu_sigma, lr scheduler)
    TValue_Trainer.init_TValue_function / train_rollout   TVT:180-248 (imported as a module) -> tvalue_trainer.npz

They run on a stand-in ``self`` carrying exactly what they read.  What stays THIRD PARTY (rl_games==1.5.2, requirements.txt:6,
not in the tree, not installed): ``discount_values`` (GAE), ``common_losses.actor_loss / critic_loss``, ``bound_loss``,
``torch_ext.policy_kl``, the ``AdaptiveScheduler`` and ``swap_and_flatten01``.  Their published algorithms are restated below
(``RLG_*``) and handed to the reference code as the collaborators it calls; the goldens record what the REFERENCE passes to
them and, for completeness, what they return.
"""
import ast
import os
import random
import sys
import tempfile
import types
from unittest import mock

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as G  # noqa: E402

REF = G.REF
OUT = G.OUT
RGC = os.path.join(REF, "utils", "rl_games_custom.py")


# ------------------------------------------------------------------ rl_games 1.5.2, published algorithms (third party)
def RLG_swap_and_flatten01(arr):
    """rl_games.common.a2c_common.swap_and_flatten01: [H, N, ...] -> [N * H, ...], env-major"""
    if arr is None:
        return arr
    s = arr.size()
    return arr.transpose(0, 1).reshape(s[0] * s[1], *s[2:])


def RLG_discount_values(self, fdones, last_extrinsic_values, mb_fdones, mb_extrinsic_values, mb_rewards):
    """rl_games.common.a2c_common.A2CBase.discount_values (GAE)"""
    lastgaelam = 0
    mb_advs = torch.zeros_like(mb_rewards)
    for t in reversed(range(self.horizon_length)):
        if t == self.horizon_length - 1:
            nextnonterminal = 1.0 - fdones
            nextvalues = last_extrinsic_values
        else:
            nextnonterminal = 1.0 - mb_fdones[t + 1]
            nextvalues = mb_extrinsic_values[t + 1]
        nextnonterminal = nextnonterminal.unsqueeze(1)
        delta = mb_rewards[t] + self.gamma * nextvalues * nextnonterminal - mb_extrinsic_values[t]
        mb_advs[t] = lastgaelam = delta + self.gamma * self.tau * nextnonterminal * lastgaelam
    return mb_advs


class RLG_AdaptiveScheduler:
    """rl_games.common.schedulers.AdaptiveScheduler"""

    def __init__(self, kl_threshold=0.008):
        self.min_lr, self.max_lr, self.kl_threshold = 1e-6, 1e-2, kl_threshold

    def update(self, current_lr, entropy_coef, epoch, frames, kl_dist, **kwargs):
        lr = current_lr
        if kl_dist > (2.0 * self.kl_threshold):
            lr = max(current_lr / 1.5, self.min_lr)
        if kl_dist < (0.5 * self.kl_threshold):
            lr = min(current_lr * 1.5, self.max_lr)
        return lr, entropy_coef


def RLG_mean_list(val):
    return torch.mean(torch.stack(val))


# ------------------------------------------------------------------ reference functions from source text
def ref_functions():
    tree = ast.parse(open(RGC).read(), RGC)
    ns = {"torch": torch, "np": np, "LOG2PI": np.log(2.0 * np.pi), "swap_and_flatten01": RLG_swap_and_flatten01,
          "time": __import__("time"), "torch_ext": types.SimpleNamespace(mean_list=RLG_mean_list)}
    out = {}

    def take(node, name):
        node.decorator_list = []
        mod = ast.Module(body=[node], type_ignores=[])
        ast.fix_missing_locations(mod)
        exec(compile(mod, RGC, "exec"), ns)
        out[name] = ns[node.name]

    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("_calc_neglogp", "_calc_ac_loss"):
            take(node, node.name)
        if isinstance(node, ast.ClassDef) and node.name == "A2CControllerAgent":
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name in ("play_steps", "prepare_dataset", "train_epoch"):
                    take(sub, sub.name)
    return out


class Buffer:
    """the rl_games ExperienceBuffer surface play_steps uses: update_data / tensor_dict / get_transformed_list"""

    def __init__(self):
        self.tensor_dict, self.log = {}, []

    def update_data(self, name, index, val):
        self.log.append((name, int(index)))
        if isinstance(val, mock.MagicMock) or val is None:
            return
        if name not in self.tensor_dict:
            self.tensor_dict[name] = torch.zeros((8,) + tuple(val.shape), dtype=val.dtype)
        self.tensor_dict[name][index] = val

    def get_transformed_list(self, fn, names):
        return {k: fn(self.tensor_dict[k]) for k in names if k in self.tensor_dict}


def gen_play_steps(F):
    """a scripted env and a scripted policy: the golden is WHAT IS STORED AT WHICH STEP and what reaches discount_values"""
    torch.manual_seed(7)
    H, N, A, OD, SD = 8, 8, 23, 5, 4
    obs_stream = torch.randn(H + 1, N, OD)            # obs_stream[t] is what the policy sees at step t
    st_stream = torch.randn(H + 1, N, SD)
    rew_stream = torch.randn(H, N, 1)
    done_stream = (torch.rand(H, N) < 0.3).to(torch.uint8)
    dones0 = (torch.rand(N) < 0.5).to(torch.uint8)    # dones carried in from the previous rollout
    val_stream = torch.randn(H + 1, N, 1)             # scripted critic: value of obs_stream[t]
    act_stream, nlp_stream, mu_stream = torch.randn(H, N, A), torch.randn(H, N), torch.randn(H, N, A)
    s = G.Fake()
    s.steps_num = s.horizon_length = H
    s.use_action_masks, s.has_central_value, s.num_agents = False, True, 1
    s.update_list = ["actions", "neglogpacs", "values", "mus", "sigmas"]
    s.tensor_list = s.update_list + ["obses", "states", "dones"]
    s.experience_buffer = Buffer()
    s.control_dict_converter = mock.MagicMock()
    s.model, s.algo_observer = mock.MagicMock(), mock.MagicMock()
    s.game_rewards, s.game_lengths = mock.MagicMock(), mock.MagicMock()
    s.central_value_net = types.SimpleNamespace(use_joint_obs_actions=False)
    s.control_dict, s._controls = None, None
    s.gamma, s.tau, s.batch_size = 0.99, 0.95, H * N
    s.current_rewards, s.current_lengths = torch.zeros(N, 1), torch.zeros(N)
    s.rewards_shaper = lambda r: r * 1.0              # reward_shaper scale_value 1 (yaml)
    s.postprocess_obs = lambda o: o
    s.obs = {"obs": obs_stream[0], "states": st_stream[0], "control_dict": None}
    s.dones = dones0.clone()
    step = {"t": 0}

    def get_action_values(obs):
        t = step["t"]
        assert torch.equal(obs["obs"], obs_stream[t])
        return {"actions": act_stream[t], "neglogpacs": nlp_stream[t], "values": val_stream[t], "mus": mu_stream[t],
                "sigmas": torch.ones(N, A)}

    def env_step(actions):
        t = step["t"]
        assert torch.equal(actions, act_stream[t])
        step["t"] = t + 1
        return ({"obs": obs_stream[t + 1], "states": st_stream[t + 1], "control_dict": None}, rew_stream[t], done_stream[t].clone(), {})

    gae_args = {}

    def discount_values(fdones, last_values, mb_fdones, mb_values, mb_rewards):
        gae_args.update(fdones=fdones.clone(), last_values=last_values.clone(), mb_fdones=mb_fdones.clone(), mb_values=mb_values.clone(),
                        mb_rewards=mb_rewards.clone())
        gae_args["advs"] = RLG_discount_values(s, fdones, last_values, mb_fdones, mb_values, mb_rewards)
        return gae_args["advs"]

    s.get_action_values, s.env_step, s.discount_values = get_action_values, env_step, discount_values
    s.get_values = lambda obs: val_stream[H] if torch.equal(obs["obs"], obs_stream[H]) else None
    batch = F["play_steps"](s)
    advs = gae_args["advs"]
    np.savez(os.path.join(OUT, "ppo_play_steps.npz"),
             obs_stream=obs_stream.numpy(), st_stream=st_stream.numpy(), rew_stream=rew_stream.numpy(), done_stream=done_stream.numpy(),
             dones0=dones0.numpy(), val_stream=val_stream.numpy(), act_stream=act_stream.numpy(), nlp_stream=nlp_stream.numpy(),
             mu_stream=mu_stream.numpy(),
             gae_fdones=gae_args["fdones"].numpy(), gae_last_values=gae_args["last_values"].numpy(), gae_mb_fdones=gae_args["mb_fdones"].numpy(),
             gae_mb_values=gae_args["mb_values"].numpy(), gae_mb_rewards=gae_args["mb_rewards"].numpy(), gae_advs=advs.numpy(),
             batch_obses=batch["obses"].numpy(), batch_states=batch["states"].numpy(), batch_dones=batch["dones"].numpy(),
             batch_values=batch["values"].numpy(), batch_returns=batch["returns"].numpy(), batch_actions=batch["actions"].numpy(),
             batch_neglogpacs=batch["neglogpacs"].numpy(),
             store_order=np.array([f"{n}:{i}" for n, i in s.experience_buffer.log if i < 2]))
    return batch


def gen_prepare_dataset(F):
    torch.manual_seed(8)
    B, A = 96, 23
    bd = {k: None for k in ("obses", "next_obses", "control_dicts", "next_control_dicts", "control_goals", "controls", "dones", "actions",
                            "pre_actions", "neglogpacs", "mus", "sigmas", "states")}
    bd["returns"], bd["values"] = torch.randn(B, 1) * 2 + 1, torch.randn(B, 1)
    s = G.Fake()
    s.normalize_value, s.normalize_advantage, s.is_rnn, s.has_central_value = False, True, False, True
    got = {}
    s.dataset = types.SimpleNamespace(update_values_dict=lambda d: got.update(actor=d))
    s.central_value_net = types.SimpleNamespace(update_dataset=lambda d: got.update(cv=d))
    F["prepare_dataset"](s, bd)
    np.savez(os.path.join(OUT, "ppo_prepare_dataset.npz"), returns=bd["returns"].numpy(), values=bd["values"].numpy(),
             advantages=got["actor"]["advantages"].numpy(), cv_advantages=got["cv"]["advantages"].numpy(),
             old_values=got["actor"]["old_values"].numpy(), cv_returns=got["cv"]["returns"].numpy())


def gen_schedule(F):
    """train_epoch's loop: minibatch order, update_mu_sigma after EVERY minibatch, and -- with the schedule_type the SeqDex yamls
    leave at rl_games' default 'legacy' -- the lr scheduler after every minibatch too (RGC:1360-1365)"""
    rng = np.random.default_rng(9)
    mini_epochs, nmb = 5, 4
    kls = rng.uniform(0.0, 0.08, size=mini_epochs * nmb).astype(np.float32)
    for sched in ("legacy", "standard"):
        s = G.Fake()
        s.set_eval = s.set_train = lambda: None
        s.is_rnn, s.has_central_value, s.has_phasic_policy_gradients, s.multi_gpu = False, True, False, False
        s.play_steps = lambda: {"played_frames": 64}
        s.prepare_dataset = lambda b: None
        s.algo_observer, s.model = mock.MagicMock(), mock.MagicMock()
        events = []
        s.train_central_value = lambda: events.append("cv")
        s.mini_epochs_num, s.bounds_loss_coef, s.schedule_type = mini_epochs, 0.001, sched
        s.scheduler = RLG_AdaptiveScheduler(0.02)                    # kl_threshold 0.02 (cfg/lego/ppo_continuous_grasp.yaml)
        s.last_lr, s.entropy_coef, s.epoch_num = 3e-4, 0.0, 0
        k = {"i": 0}

        class DS:
            def __len__(self):
                return nmb

            def __getitem__(self, i):
                return i

            def update_mu_sigma(self, mu, sigma):
                events.append(f"mu_sigma:{mu}")
        s.dataset = DS()
        lrs = []

        def train_actor_critic(i):
            kl = torch.tensor(float(kls[k["i"]]))
            k["i"] += 1
            events.append(f"mb:{i}")
            z = torch.zeros(())
            return z, z, z, kl, s.last_lr, 1.0, i, i, z, {}
        s.train_actor_critic = train_actor_critic
        s.update_lr = lambda lr: (lrs.append(lr), events.append("lr"))
        F["train_epoch"](s)
        np.savez(os.path.join(OUT, f"ppo_schedule_{sched}.npz"), kls=kls, lrs=np.asarray(lrs, np.float64), events=np.array(events),
                 mini_epochs=mini_epochs, nmb=nmb, lr0=3e-4, kl_threshold=0.02)


def gen_tvalue_trainer():
    """TVT:180-248 on CPU tensors: one call of train_rollout with rollout = 1 (batch 512 + 512, noise, re-normalisation, BCE, Adam)"""
    import policy_sequencing.transition_value_trainer as TVT
    sys.path.insert(0, os.path.dirname(REF))           # TVT:181 imports dexteroushandenvs.policy_sequencing...
    torch.manual_seed(11)
    random.seed(11)
    rng = np.random.default_rng(11)
    q = rng.normal(size=(4000, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    succ, fail = torch.from_numpy(q[q[:, 3] > 0.1][:1500]), torch.from_numpy(q[q[:, 3] < -0.1][:1500])
    t = object.__new__(TVT.TValue_Trainer)
    t.device, t.input_dim = "cpu", 4
    t.valid_data, t.success_data, t.failure_data = succ[-100:].clone(), succ[:-100].clone(), fail.clone()
    t.num_success_data, t.num_failure_data = t.success_data.shape[0], t.failure_data.shape[0]
    losses = []
    TVT.print = lambda *a: losses.append(a[1]) if a and a[0] == "loss: " else None
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            t.init_TValue_function("golden", 1)
            w0 = torch.cat([p.detach().reshape(-1) for p in t.t_value.parameters()]).clone()
            t.train_rollout()
        finally:
            os.chdir(cwd)
    w1 = torch.cat([p.detach().reshape(-1) for p in t.t_value.parameters()])
    np.savez(os.path.join(OUT, "tvalue_trainer.npz"), success_data=t.success_data.numpy(), failure_data=t.failure_data.numpy(),
             rand_float=t.rand_float.numpy(), succ_rand=np.asarray(t.succ_rand, np.int64), fail_rand=np.asarray(t.fail_rand, np.int64),
             obs_buf=t.t_value_obs_buf.numpy(), target=t.success_buf.numpy(), logits=t.predict_success_confident.detach().numpy(),
             loss=np.float32(losses[0]), w0=w0.numpy(), w1=w1.numpy())


def main():
    os.makedirs(OUT, exist_ok=True)
    G.install_stubs()
    F = ref_functions()
    # ---- _calc_neglogp (RGC:2113-2127), the non-tanh branch the SeqDex yamls use
    torch.manual_seed(5)
    M, A = 64, 23
    x, mean, logstd = torch.randn(M, A), torch.randn(M, A) * 0.5, torch.randn(A) * 0.2
    std = torch.exp(logstd).expand(M, A)
    nlp = F["_calc_neglogp"](x, x, mean, std, logstd.expand(M, A), False)
    np.savez(os.path.join(OUT, "ppo_neglogp.npz"), x=x.numpy(), mean=mean.numpy(), logstd=logstd.numpy(), neglogp=nlp.numpy())
    # ---- _calc_ac_loss (RGC:2129-2132)
    v = torch.randn(16, 4)
    loss = torch.stack([F["_calc_ac_loss"](a, c, 1.0, e, 0.0, b, 0.001) for a, c, e, b in v])
    loss4 = torch.stack([F["_calc_ac_loss"](a, c, 4.0, e, 0.0, b, 0.001) for a, c, e, b in v])      # ppo_continuous_insert.yaml critic_coef 4
    np.savez(os.path.join(OUT, "ppo_ac_loss.npz"), terms=v.numpy(), loss=loss.numpy(), loss_critic_coef4=loss4.numpy())
    gen_play_steps(F)
    gen_prepare_dataset(F)
    gen_schedule(F)
    gen_tvalue_trainer()
    print("wrote ppo goldens to", OUT)


if __name__ == "__main__":
    main()
