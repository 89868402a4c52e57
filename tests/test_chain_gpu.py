"""BASELINE configs[3]: Search -> Orient -> GraspSim -> InsertSim on one GPU with the hand-offs on the device, and the reference's
pickles written beside them (seqdex_b200/chain.py)."""
import os
import pickle

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_chain_search_orient_grasp_insert(tmp_path):
    from seqdex_b200.chain import run_chain
    from seqdex_b200.tasks.block_assembly_grasp_sim import default_tvalue_weights
    w = default_tvalue_weights(1)
    w[-1] += 50.0                                   # gates open: this test is about the plumbing, not about a trained gate
    out = run_chain(num_envs=64, episodes=(1, 1, 1, 1), tvalue_weights=w, bank_capacity=16, save_dir=str(tmp_path))
    assert out["search_heaps_per_type"] >= 1 and out["orient_heaps_per_type"] >= 1
    for k in ("search_mean_reward", "orient_mean_reward", "grasp_mean_reward"):
        assert out[k] == out[k] and abs(out[k]) < 1e4, (k, out[k])
    assert 0.0 < out["orient_mean_reward"] <= 1.0
    names = ["saved_searching_ternimal_states_medium_mo_tvalue.pkl", "saved_searching_hand_ternimal_states_medium_mo_tvalue.pkl",
             "saved_searching_ternimal_states_good_mo_tvalue.pkl", "saved_grasping_hand_ternimal_states_good_mo_sim.pkl",
             "saved_grasping_object_ternimal_states_good_mo_sim.pkl"]
    shapes = [(11024, 132, 13), (11024, 23, 2), (11024, 132, 13), (11024, 23, 2), (11024, 1, 13)]
    for name, shp in zip(names, shapes):
        with open(os.path.join(tmp_path, name), "rb") as f:
            lst = pickle.load(f)
        assert isinstance(lst, list) and len(lst) == 8 and all(tuple(t.shape) == shp for t in lst), name
    s, o = out["banks"]["search"], out["banks"]["orient"]
    assert s.shape[2:] == (72, 13) and o.shape[2:] == (72, 13) and torch.isfinite(s).all() and torch.isfinite(o).all()
    # the heaps Orient hands on went through two stages of random pushing: bricks rest in the bin (a few may have been
    # knocked over its wall onto the table or the ground, and a snapshot can catch the odd brick in the air: the hand flails
    # at full speed and max_lin_vel is Isaac Gym's 1000 m/s), none is below the ground, nothing has exploded
    # (physically: a brick pinched between two closing links is ejected like a pip; a 3.8 m/s launch reaches 1.4 m)
    z, x, y = o[..., 2].flatten(), o[..., 0].flatten(), o[..., 1].flatten()
    assert float(z.min()) > 0.013                                   # nothing below the ground: a brick lying on it has its origin >= 1.5 cm up
    assert float(z.max()) < 2.0 and 0.6 < float(z.median()) < 0.8    # nothing has exploded
    in_bin = (x > -0.06) & (x < 0.56) & (y > -0.03) & (y < 0.41) & (z > 0.60) & (z < 1.0)
    assert float(in_bin.float().mean()) > 0.95, float(in_bin.float().mean())     # >= 95 % of the banked bricks rest inside the bin (measured 96.7 %: a
    # random policy's hand ploughs through the heap at full speed for two episodes and Orient's script lifts it 42 cm with bricks on it)
    # ---- stage 4: InsertSim consumed the grasp rings (or, for brick types an untrained policy never lifted, the synthetic stand-ins)
    assert out["insert_bank_rows_per_type"] >= 1 and 0 <= out["insert_bank_synthetic_types"] <= 8
    assert out["insert_mean_reward"] == out["insert_mean_reward"] and 0.0 <= out["insert_mean_reward"] <= 2.0     # bonus + exp(-...) <= 2 (IS:1672)
    assert 0.0 <= out["insert_success_rate"] <= 1.0
    gh, go = out["banks"]["grasp"]
    assert gh.shape[0] == 8 and gh.shape[2:] == (23, 2) and go.shape[2:] == (13,) and torch.isfinite(gh).all() and torch.isfinite(go).all()
