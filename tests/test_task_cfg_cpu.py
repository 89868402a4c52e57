"""A task runs with the configuration its yaml states (Orient / Search ran with a 10x smaller contact_offset).
tests/golden/task_cfg.json = the `env:` scalars and the whole `sim:` block of the reference yamls (oracle/gen_golden_cfg.py);
seqdex_b200/tasks/cfg.py must equal it field for field, and the Scene every entry point builds must carry those values."""
import json
import os

import pytest

from seqdex_b200.tasks.cfg import TASK_CFG, scene_from_cfg

FIX = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "task_cfg.json")))


@pytest.mark.parametrize("task", sorted(FIX))
def test_default_cfg_equals_the_reference_yaml(task):
    assert TASK_CFG[task]["env"] == FIX[task]["env"]
    assert TASK_CFG[task]["sim"] == FIX[task]["sim"]
    assert TASK_CFG[task]["task"] == FIX[task]["task"]


@pytest.mark.parametrize("task", ["BlockAssemblySearch", "BlockAssemblyOrient", "BlockAssemblyGraspSim"])
def test_scene_carries_the_yaml_values(task):
    y = FIX[task]
    sc = scene_from_cfg(task)
    c = sc.c
    assert c.substeps == y["sim"]["substeps"] and c.iters == y["sim"]["physx"]["num_position_iterations"]
    assert abs(c.contact_offset - y["sim"]["physx"]["contact_offset"]) < 1e-9
    assert c.max_episode_length == y["env"]["episodeLength"]
    assert abs(c.act_moving_average - y["env"]["actionsMovingAverage"]) < 1e-7
    assert sc.yaml_max_depen_vel == y["sim"]["physx"]["max_depenetration_velocity"]
    assert abs(c.max_depen_vel - min(sc.yaml_max_depen_vel, sc.push_out_cap)) < 1e-7      # the solver's own cap is explicit, not a silent edit
    assert abs(c.dt - 1.0 / 60.0) < 1e-9 and abs(c.gravity_z + 9.81) < 1e-6
    assert abs(c.face_margin - 0.002) < 1e-9                     # in-face tolerance does NOT grow with the speculative offset
    # a partial cfg (what tests / chain.py pass) falls back to the yaml, not to another task's numbers
    sc2 = scene_from_cfg(task, {"env": {"numEnvs": 8}, "sim": {"physx": {}}})
    assert abs(sc2.c.contact_offset - y["sim"]["physx"]["contact_offset"]) < 1e-9 and sc2.c.max_episode_length == y["env"]["episodeLength"]


def test_the_three_tasks_differ_where_the_yamls_differ():
    assert abs(scene_from_cfg("BlockAssemblyOrient").c.contact_offset - 0.02) < 1e-9
    assert abs(scene_from_cfg("BlockAssemblySearch").c.contact_offset - 0.02) < 1e-9
    assert abs(scene_from_cfg("BlockAssemblyGraspSim").c.contact_offset - 0.002) < 1e-9
