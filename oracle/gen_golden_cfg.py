#!/usr/bin/env python
"""The `env:` scalars and the whole `sim:` block of the reference's task yamls, as a committed fixture (tests/golden/task_cfg.json).
Build container only (reads /root/reference).  tests/test_task_cfg_cpu.py holds every task's DEFAULT_CFG (seqdex_b200/tasks/cfg.py)
and the Scene it builds against this fixture field for field -- a task must run with the configuration its yaml states."""
import json
import os

import yaml

REF = "/root/reference/dexteroushandenvs/cfg"
OUT = os.environ.get("SEQDEX_GOLDEN_OUT") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
YAMLS = {
    "BlockAssemblySearch": "allegro_hand_block_assembly_search.yaml",
    "BlockAssemblyOrient": "allegro_hand_block_assembly_orient.yaml",
    "BlockAssemblyGraspSim": "allegro_hand_block_assembly_grasp_sim.yaml",
    "BlockAssemblyInsertSim": "allegro_hand_block_assembly_insert_sim.yaml",
    "ToolPositioningGrasp": "allegro_hand_tool_positioning_grasp.yaml",
    "ToolPositioningOrient": "allegro_hand_tool_positioning_orient.yaml",
}


def main():
    out = {}
    for task, f in YAMLS.items():
        d = yaml.safe_load(open(os.path.join(REF, f)))
        out[task] = {"yaml": f, "env": {k: v for k, v in d["env"].items() if not isinstance(v, (dict, list))}, "sim": d["sim"],
                     "task": {"randomize": d["task"]["randomize"]}}
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "task_cfg.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("wrote", os.path.join(OUT, "task_cfg.json"))


if __name__ == "__main__":
    main()
