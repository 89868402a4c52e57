#!/usr/bin/env python
"""the physical summary tests/test_chain_gpu.py asserts on, for whichever library SEQDEX_B200_LIB names (A/B of builds)"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seqdex_b200.chain import run_chain                                    # noqa: E402
from seqdex_b200.tasks.block_assembly_grasp_sim import default_tvalue_weights   # noqa: E402

w = default_tvalue_weights(1)
w[-1] += 50.0
with tempfile.TemporaryDirectory() as d:
    out = run_chain(num_envs=64, episodes=(1, 1, 1), tvalue_weights=w, bank_capacity=16, save_dir=d)
o, s = out["banks"]["orient"], out["banks"]["search"]
print(os.environ.get("SEQDEX_B200_LIB", "default"), "orient z min/max/median", float(o[..., 2].min()), float(o[..., 2].max()), float(o[..., 2].median()),
      "search z max", float(s[..., 2].max()), "heaps", out["search_heaps_per_type"], out["orient_heaps_per_type"],
      "rewards", out["search_mean_reward"], out["orient_mean_reward"], out["grasp_mean_reward"])
