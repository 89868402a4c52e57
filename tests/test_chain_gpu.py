"""BASELINE configs[3], the links that exist: Search -> Orient -> GraspSim on one GPU with the hand-offs on the device, and the
reference's pickles written beside them (seqdex_b200/chain.py)."""
import os
import pickle

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_chain_search_orient_grasp(tmp_path):
    from seqdex_b200.chain import run_chain
    from seqdex_b200.tasks.block_assembly_grasp_sim import default_tvalue_weights
    w = default_tvalue_weights(1)
    w[-1] += 50.0                                   # gates open: this test is about the plumbing, not about a trained gate
    out = run_chain(num_envs=64, episodes=(1, 1, 1), tvalue_weights=w, bank_capacity=16, save_dir=str(tmp_path))
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
    z = o[..., 2].flatten()
    assert float(z.min()) > -0.02 and float(z.abs().max()) < 100.0 and 0.6 < float(z.median()) < 0.8
    assert float((z > 1.2).float().mean()) < 0.01
