"""The committed golden vectors ARE what the reference's own Python produces: where the reference tree is present (the build
container; not the GPU box) every generator under oracle/ is re-run into a scratch directory and its output compared with
tests/golden/ bit for bit.  (The generators execute reference functions on stand-in objects; see their docstrings.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GENERATORS = {
    "gen_golden.py": ["tvalue.npz", "control_ik.npz", "post_physics.npz", "pre_physics.npz"],
    "gen_golden_orient.py": ["orient_post_physics.npz", "orient_pre_physics.npz"],
    "gen_golden_search.py": ["search_post_physics.npz", "search_pre_physics.npz"],
    "gen_golden_dr.py": ["dr_params.npz"],
    "gen_golden_reset.py": ["reset_idx.npz"],
    "gen_golden_cfg.py": ["task_cfg.json"],
    "gen_golden_insert.py": ["insert_post_physics.npz", "insert_pre_physics.npz", "insert_reset.npz"],
    "gen_golden_tool.py": ["tool_grasp_post.npz", "tool_grasp_pre.npz", "tool_grasp_reset.npz", "tool_orient_post.npz", "tool_orient_pre.npz",
                           "tool_orient_reset.npz"],
    "gen_golden_ppo.py": ["ppo_neglogp.npz", "ppo_ac_loss.npz", "ppo_play_steps.npz", "ppo_prepare_dataset.npz", "ppo_schedule_legacy.npz",
                          "ppo_schedule_standard.npz", "tvalue_trainer.npz"],
}

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/dexteroushandenvs"), reason="needs the reference tree (build container only)")


@pytest.mark.parametrize("script", sorted(GENERATORS))
def test_generator_reproduces_the_committed_vectors(script, tmp_path):
    env = dict(os.environ, SEQDEX_GOLDEN_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    for name in GENERATORS[script]:
        if name.endswith(".json"):
            assert open(os.path.join(tmp_path, name)).read() == open(os.path.join(GOLDEN, name)).read(), name
            continue
        new, old = np.load(os.path.join(tmp_path, name)), np.load(os.path.join(GOLDEN, name))
        assert set(new.files) == set(old.files), name
        for k in new.files:
            assert np.array_equal(new[k], old[k]), f"{name}:{k} differs from the committed golden vector"


def test_every_golden_file_has_a_generator():
    made = {n for names in GENERATORS.values() for n in names} | {"dr_configs.json"}
    made |= {"facade_dump.npz"}     # produced ON A B200 by tools/dump_facade.py (the facade's tensors); tests/test_facade_gpu.py re-creates it bit for bit
    assert set(os.listdir(GOLDEN)) == made
