"""Transition-feasibility (t-value) trainer with the reference's semantics
(policy_sequencing/transition_value_trainer.py TVT:127-248): GraspInsertTValue 4 -> 256 -> 128 -> 64 -> 2 (ELU after
every layer incl. the last, TVF:30-46) trained with BCE-with-logits against one-hot [failure, success], Adam 1e-3,
batch = 512 successes + 512 failures with U(-0.05, 0.05) noise and quaternion re-normalisation (TVT:209-226).
Forward/backward/Adam run on the same tensor-core MLP object as PPO; the loss is one fused kernel.  The trained weights
plug into the env's gate kernel through ``SdxEnv.set_tvalue_weights`` (GS:1200-1201, 1406)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .ppo import MLP, _p, _stream


class TValueTrainer:
    def __init__(self, success, failure, device=0, seed=0, batch=1024, lr=1e-3, holdout=100):
        """success / failure: [n, 4] camera-frame target quaternions (the HDF5 '{i}th_success_data' rows, TVT:132-168)"""
        self.device = torch.device("cuda", device)
        g = torch.Generator().manual_seed(seed)
        s = torch.as_tensor(success, dtype=torch.float32)
        f = torch.as_tensor(failure, dtype=torch.float32)
        s, f = s[torch.randperm(len(s), generator=g)], f[torch.randperm(len(f), generator=g)]
        holdout = min(holdout, len(s) // 4, len(f) // 4)
        self.val_x = torch.cat([s[:holdout], f[:holdout]]).to(self.device)                  # TVT:172-173
        self.val_y = torch.cat([torch.ones(holdout), torch.zeros(holdout)]).to(self.device)
        self.succ, self.fail = s[holdout:].to(self.device), f[holdout:].to(self.device)
        self.batch, self.lr = batch, lr
        self.net = MLP(4, 2, batch, device=device, seed=seed, hidden=(256, 128, 64))
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self.dz = torch.zeros(batch, 2, device=self.device)
        self.stats = torch.zeros(4, device=self.device)
        self.L = _lib.load()

    @staticmethod
    def make_batch(succ_rows, fail_rows, noise_s, noise_f):
        """TVT:216-220: rows + 0.05 x U(-1, 1) noise, each half re-normalised to a unit quaternion; labels 1 (success) then 0"""
        xs = succ_rows + noise_s * 0.05
        xs = xs / xs.norm(dim=-1, keepdim=True)
        xf = fail_rows + noise_f * 0.05
        xf = xf / xf.norm(dim=-1, keepdim=True)
        h = xs.shape[0]
        y = torch.cat([torch.ones(h, dtype=torch.int32, device=xs.device), torch.zeros(xf.shape[0], dtype=torch.int32, device=xs.device)])
        return torch.cat([xs, xf]).contiguous(), y

    def _sample(self):
        h = self.batch // 2
        i = torch.randint(len(self.succ), (h,), device=self.device, generator=self.gen)
        j = torch.randint(len(self.fail), (h,), device=self.device, generator=self.gen)
        noise = torch.rand(2 * h, 4, device=self.device, generator=self.gen) * 2 - 1         # TVT:210 torch_rand_float(-1, 1, ...)
        return self.make_batch(self.succ[i], self.fail[j], noise[:h], noise[h:])

    def step(self, x, y):
        """forward, BCE-with-logits on the ELU outputs, backward, Adam(lr) (TVT:222-229); returns the device statistics (sum of losses at [0])"""
        z = self.net.forward(x, train=True)
        self.stats.zero_()
        _lib.check(self.L.sdx_tvalue_bce(_p(z), _p(y), x.shape[0], _p(self.dz), _p(self.stats), _stream()))
        self.net.backward(self.dz)
        self.net.adam(self.lr, max_norm=0.0)
        return self.stats

    def train_rollout(self, iters):
        last = None
        for _ in range(iters):
            last = self.step(*self._sample())
        return float(last[0]) / (2 * self.batch) if last is not None else float("nan")

    @torch.no_grad()
    def validate(self):
        z = self.net.forward(self.val_x.contiguous())
        y = torch.nn.functional.elu(z)
        pred = torch.sigmoid(y)[:, 1] > 0.5                    # the gate reads sigmoid(.)[:, 1] (GS:1201)
        return float((pred.float() == self.val_y).float().mean())

    def weights(self):
        """flat state_dict-order weights for SdxEnv.set_tvalue_weights (W1 b1 W2 b2 W3 b3 W4 b4)"""
        return self.net.params.detach().cpu().numpy().copy()
