"""CPU restatement (plain torch fp32) of the rl_games half of the hot path -- SURVEY.md section 8 rows a13-a16.  TEST INFRASTRUCTURE
ONLY: imported by tests/ (and nothing else).

Pinned by tests/test_ppo_oracle_golden.py against vectors produced by EXECUTING the reference's own code
(oracle/gen_golden_ppo.py -> tests/golden/ppo_*.npz, tvalue_trainer.npz): ``_calc_neglogp`` RGC:2113-2127, ``_calc_ac_loss``
RGC:2129-2132, ``play_steps`` RGC:1394-1483, ``prepare_dataset`` RGC:1621-1683, ``train_epoch`` RGC:1306-1392, ``TValue_Trainer``
TVT:180-248.  What the reference delegates to rl_games==1.5.2 (absent; requirements.txt:6) is restated from its published source
and marked THIRD PARTY: those functions are anchored on the reference's call sites, not on reference-executed vectors."""
import math

import torch


def neglogp(x, mean, logstd):
    """RGC:2113-2127, use_tanh False"""
    return 0.5 * (((x - mean) / torch.exp(logstd)) ** 2).sum(-1) + 0.5 * math.log(2.0 * math.pi) * x.shape[-1] + logstd.sum(-1)


def ac_loss(a_loss, c_loss, critic_coef, entropy, entropy_coef, b_loss, bounds_loss_coef):
    """RGC:2129-2132"""
    return a_loss + 0.5 * c_loss * critic_coef - entropy * entropy_coef + b_loss * bounds_loss_coef


def actor_loss(old_neglogp, new_neglogp, adv, e_clip):
    """THIRD PARTY rl_games.common.common_losses.actor_loss (ppo=True); call site RGC:1814"""
    ratio = torch.exp(old_neglogp - new_neglogp)
    return torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - e_clip, 1.0 + e_clip))


def critic_loss(old_values, values, e_clip, returns, clip_value=True):
    """THIRD PARTY rl_games.common.common_losses.critic_loss; call site RGC:1818"""
    if clip_value:
        vc = old_values + (values - old_values).clamp(-e_clip, e_clip)
        return torch.max((values - returns) ** 2, (vc - returns) ** 2)
    return (returns - values) ** 2


def bound_loss(mu, soft_bound=1.1):
    """THIRD PARTY rl_games ContinuousA2CBase.bound_loss; call site RGC:1823"""
    return (torch.clamp_min(mu - soft_bound, 0.0) ** 2 + torch.clamp_max(mu + soft_bound, 0.0) ** 2).sum(-1)


def policy_kl(p0_mu, p0_sigma, p1_mu, p1_sigma):
    """THIRD PARTY rl_games.algos_torch.torch_ext.policy_kl, per sample; called as policy_kl(mu, sigma, old_mu, old_sigma) at RGC:1903"""
    c1 = torch.log(p1_sigma / p0_sigma + 1e-5)
    c2 = (p0_sigma ** 2 + (p1_mu - p0_mu) ** 2) / (2.0 * (p1_sigma ** 2 + 1e-5))
    return (c1 + c2 - 0.5).sum(-1)


def normalize_advantages(returns, values):
    """RGC:1640, 1651 (torch.std is the unbiased estimator)"""
    adv = (returns - values).sum(dim=1)
    return (adv - adv.mean()) / (adv.std() + 1e-8)


def swap_and_flatten01(t):
    """THIRD PARTY rl_games.common.a2c_common.swap_and_flatten01; call sites RGC:1480-1481"""
    s = t.size()
    return t.transpose(0, 1).reshape(s[0] * s[1], *s[2:])


def tvalue_forward(w, x):
    """GraspInsertTValue (TVF:30-46) on a flat state_dict-order weight vector: ELU after every layer including the last"""
    off, h = 0, x
    for o, i in ((256, 4), (128, 256), (64, 128), (2, 64)):
        W = w[off:off + o * i].view(o, i); off += o * i
        b = w[off:off + o]; off += o
        h = torch.nn.functional.elu(h @ W.T + b)
    return h


def tvalue_batch(success, failure, succ_idx, fail_idx, rand_float):
    """TVT:210-220"""
    xs = success[succ_idx] + rand_float[:, 0:4] * 0.05
    xs = xs / xs.norm(dim=-1, keepdim=True)
    xf = failure[fail_idx] + rand_float[:, 4:8] * 0.05
    xf = xf / xf.norm(dim=-1, keepdim=True)
    return torch.cat([xs, xf])


def tvalue_loss(logits, target):
    """TVT:199, 226: BCEWithLogitsLoss (mean over batch x 2) on the network's ELU outputs"""
    return torch.nn.functional.binary_cross_entropy_with_logits(logits, target)
