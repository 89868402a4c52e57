"""Rows a13-a16 of SURVEY.md section 8: the CPU restatements the GPU tests compare against (oracle/ppo_oracle.py, the C oracle's GAE,
ppo.adaptive_lr) are pinned here to vectors produced by EXECUTING the reference's own Python (oracle/gen_golden_ppo.py)."""
import ctypes
import os

import numpy as np
import torch

from oracle import ppo_oracle as PO

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
T = torch.from_numpy


def test_neglogp_matches_the_reference_function():
    g = np.load(os.path.join(G, "ppo_neglogp.npz"))
    out = PO.neglogp(T(g["x"]), T(g["mean"]), T(g["logstd"]))
    torch.testing.assert_close(out, T(g["neglogp"]), rtol=1e-6, atol=1e-5)


def test_ac_loss_matches_the_reference_function():
    g = np.load(os.path.join(G, "ppo_ac_loss.npz"))
    t = T(g["terms"])
    torch.testing.assert_close(PO.ac_loss(t[:, 0], t[:, 1], 1.0, t[:, 2], 0.0, t[:, 3], 0.001), T(g["loss"]), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(PO.ac_loss(t[:, 0], t[:, 1], 4.0, t[:, 2], 0.0, t[:, 3], 0.001), T(g["loss_critic_coef4"]), rtol=1e-6, atol=1e-7)


def test_play_steps_bookkeeping_and_gae(oracle_lib):
    """what the reference's play_steps hands to discount_values: dones stored BEFORE step t, the rollout's last dones and the value
    of the observation AFTER the last step for the bootstrap; the C oracle's GAE sweep on exactly those arrays; env-major batch"""
    g = np.load(os.path.join(G, "ppo_play_steps.npz"))
    H, N = g["rew_stream"].shape[:2]
    # the bookkeeping, stated on the scripted streams
    pre_dones = np.concatenate([g["dones0"][None], g["done_stream"][:-1]]).astype(np.float32)
    np.testing.assert_array_equal(g["gae_mb_fdones"], pre_dones)
    np.testing.assert_array_equal(g["gae_fdones"], g["done_stream"][-1].astype(np.float32))
    np.testing.assert_array_equal(g["gae_last_values"], g["val_stream"][H])
    np.testing.assert_array_equal(g["gae_mb_values"], g["val_stream"][:H])
    np.testing.assert_array_equal(g["gae_mb_rewards"], g["rew_stream"])
    assert list(g["store_order"][:2]) == ["obses:0", "dones:0"]          # stored before the env step (RGC:1403-1404)
    # GAE: the C oracle (what k_gae is compared with bit for bit) on the reference's arguments
    L = oracle_lib.lib()
    f = lambda a: np.ascontiguousarray(a, np.float32)
    rew, val, dn = f(g["gae_mb_rewards"][..., 0]), f(g["gae_mb_values"][..., 0]), f(g["gae_mb_fdones"])
    lv, ld = f(g["gae_last_values"][..., 0]), f(g["gae_fdones"])
    adv, ret = np.zeros((H, N), np.float32), np.zeros((H, N), np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    L.sdxo_gae(P(rew), P(val), P(dn), P(lv), P(ld), P(adv), P(ret), H, N, ctypes.c_float(0.99), ctypes.c_float(0.95))
    np.testing.assert_allclose(adv, g["gae_advs"][..., 0], rtol=1e-5, atol=1e-6)
    # env-major flattening of the batch (swap_and_flatten01)
    np.testing.assert_allclose(PO.swap_and_flatten01(T(ret)).numpy(), g["batch_returns"][:, 0], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(PO.swap_and_flatten01(T(g["obs_stream"][:H])).numpy(), g["batch_obses"])
    np.testing.assert_array_equal(PO.swap_and_flatten01(T(g["act_stream"])).numpy(), g["batch_actions"])


def test_advantage_normalisation_matches_prepare_dataset():
    g = np.load(os.path.join(G, "ppo_prepare_dataset.npz"))
    out = PO.normalize_advantages(T(g["returns"]), T(g["values"]))
    torch.testing.assert_close(out, T(g["advantages"]), rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(g["advantages"], g["cv_advantages"])
    np.testing.assert_array_equal(g["old_values"], g["values"])


def test_adaptive_lr_runs_after_every_minibatch():
    """schedule_type is not set in the SeqDex yamls -> rl_games' default 'legacy': scheduler.update after EVERY minibatch"""
    from seqdex_b200.ppo import adaptive_lr
    g = np.load(os.path.join(G, "ppo_schedule_legacy.npz"))
    lr, out = float(g["lr0"]), []
    for kl in g["kls"]:
        lr = adaptive_lr(lr, float(kl), float(g["kl_threshold"]))
        out.append(lr)
    np.testing.assert_allclose(out, g["lrs"], rtol=1e-12)
    ev = list(g["events"])
    assert ev[0] == "cv"                                                # train_central_value before the actor mini-epochs (RGC:1325-1326)
    assert ev[1:4] == ["mb:0", "mu_sigma:0", "lr"]                      # minibatch, update_mu_sigma, scheduler
    assert len(g["lrs"]) == int(g["mini_epochs"]) * int(g["nmb"])
    s = np.load(os.path.join(G, "ppo_schedule_standard.npz"))
    assert len(s["lrs"]) == int(s["mini_epochs"])                       # 'standard' would have been once per mini-epoch


def test_tvalue_trainer_step_matches_the_reference():
    g = np.load(os.path.join(G, "tvalue_trainer.npz"))
    x = PO.tvalue_batch(T(g["success_data"]), T(g["failure_data"]), T(g["succ_rand"]), T(g["fail_rand"]), T(g["rand_float"]))
    torch.testing.assert_close(x, T(g["obs_buf"]), rtol=1e-6, atol=1e-7)
    w = T(g["w0"]).clone().requires_grad_(True)
    z = PO.tvalue_forward(w, x)
    torch.testing.assert_close(z.detach(), T(g["logits"]), rtol=1e-5, atol=1e-6)
    loss = PO.tvalue_loss(z, T(g["target"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    loss.backward()
    # one torch.optim.Adam(lr 1e-3) step from zero moments moves every weight by lr * sign(grad) (up to eps): compare with the reference's
    opt = torch.optim.Adam([w], lr=1e-3)
    opt.step()
    torch.testing.assert_close(w.detach(), T(g["w1"]), rtol=1e-5, atol=2e-6)
    assert np.array_equal(g["target"][:512, 1], np.ones(512)) and np.array_equal(g["target"][512:, 0], np.ones(512))
