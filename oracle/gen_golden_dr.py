#!/usr/bin/env python
"""Golden vectors for the domain-randomisation noise path, produced by EXECUTING THE REFERENCE'S OWN PYTHON
(`tasks/hand_base/base_task.py`: BaseTask.apply_randomizations BT:229-423 and the noise_lambda closures it creates, which
BaseTask.step calls at BT:131-132 and BT:149-150).  Runs only in the build container; the output is committed as
tests/golden/dr_params.npz.  Stubs as in gen_golden.py (Isaac Gym is a closed binary); everything that runs is the
reference's code on a stand-in `self`:
  * for a sweep of configurations (gaussian / uniform x additive / scaling x linear / constant / no schedule) and frame counts:
    the parameters the closure closes over (mu, var, mu_corr, var_corr | lo, hi, lo_corr, hi_corr);
  * the closure applied to a tensor with torch's generator seeded, and -- by replaying the same generator -- the very draws it
    consumed (corr = randn_like on the first call, then randn_like / rand_like per call): pins the combination formula;
  * the refresh bookkeeping over a sequence of frame counts (first_randomization, last_rand_step, frequency).
"""
import os
import sys
from unittest import mock

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import install_stubs, Fake, OUT   # noqa: E402

CONFIGS = [
    {"range": [0, .002], "range_correlated": [0, .001], "operation": "additive", "distribution": "gaussian", "schedule": "linear", "schedule_steps": 40000},
    {"range": [0., .05], "range_correlated": [0, .015], "operation": "additive", "distribution": "gaussian", "schedule": "linear", "schedule_steps": 40000},
    {"range": [1.0, .05], "range_correlated": [1.0, .02], "operation": "scaling", "distribution": "gaussian", "schedule": "linear", "schedule_steps": 3000},
    {"range": [0.1, .3], "operation": "additive", "distribution": "gaussian", "schedule": "constant", "schedule_steps": 500},
    {"range": [-0.01, .02], "range_correlated": [-0.005, .005], "operation": "additive", "distribution": "uniform", "schedule": "linear", "schedule_steps": 1000},
    {"range": [0.9, 1.1], "range_correlated": [0.95, 1.05], "operation": "scaling", "distribution": "uniform", "schedule": "linear", "schedule_steps": 2000},
    {"range": [0.9, 1.2], "operation": "scaling", "distribution": "uniform"},
]
STEPS = [0, 1, 250, 499, 500, 999, 1000, 2500, 39999, 40000, 123456]


def fresh(num_envs, step):
    f = Fake()
    f.gym, f.sim = mock.MagicMock(), None
    f.gym.get_frame_count = lambda sim: f._frame
    f._frame = step
    f.num_envs, f.envs = num_envs, [None] * num_envs
    f.first_randomization, f.last_rand_step, f.last_step = True, -1, -1
    f.dr_randomizations, f.original_props = {}, {}
    f.actor_params_generator, f.extern_actor_params = None, {i: None for i in range(num_envs)}
    f.randomize_buf = torch.zeros(num_envs, dtype=torch.long)
    f.reset_buf = torch.ones(num_envs, dtype=torch.long)
    return f


def main():
    install_stubs()
    import tasks.hand_base.base_task as BT
    out = {}
    shape = (6, 9)
    for ci, cfg in enumerate(CONFIGS):
        keys = ("mu", "var", "mu_corr", "var_corr") if cfg["distribution"] == "gaussian" else ("lo", "hi", "lo_corr", "hi_corr")
        table = np.zeros((len(STEPS), 4), np.float64)
        for si, step in enumerate(STEPS):
            f = fresh(4, step)
            BT.BaseTask.apply_randomizations(f, {"frequency": 1, "observations": dict(cfg), "actor_params": {}})
            p = f.dr_randomizations["observations"]
            table[si] = [p[k] for k in keys]
            if step in (250, 2500, 123456):
                # the closure on a tensor; replay the generator to capture the draws it consumed
                x = torch.randn(*shape, generator=torch.Generator().manual_seed(100 + ci)) * 0.7
                torch.manual_seed(777 + si)
                y1 = p["noise_lambda"](x)              # first call: draws corr, then the white noise
                y2 = p["noise_lambda"](x)              # second call: reuses corr, draws white noise again
                torch.manual_seed(777 + si)
                corr = torch.randn_like(x)
                draw = torch.randn_like if cfg["distribution"] == "gaussian" else torch.rand_like
                w1, w2 = draw(x), draw(x)
                tag = f"c{ci}_s{step}_"
                for k, v in (("x", x), ("corr", corr), ("w1", w1), ("w2", w2), ("y1", y1), ("y2", y2)):
                    out[tag + k] = v.numpy().astype(np.float32)
        out[f"c{ci}_params"] = table
    # ---- refresh bookkeeping (BT:233-249): when are the non-env parameters regenerated?
    freq = 1000
    frames = [0, 1, 500, 999, 1000, 1001, 1999, 2000, 2600, 3001, 3002, 5000]
    f = fresh(4, 0)
    cfg = dict(CONFIGS[0])
    log = []
    for fr in frames:
        f._frame = fr
        before = f.dr_randomizations.get("observations")
        BT.BaseTask.apply_randomizations(f, {"frequency": freq, "observations": dict(cfg), "actor_params": {}})
        after = f.dr_randomizations.get("observations")
        log.append([fr, int(after is not before), f.last_rand_step])
    out["refresh_log"] = np.array(log, np.int64)
    out["refresh_freq"] = np.array([freq], np.int64)
    out["steps"] = np.array(STEPS, np.int64)
    np.savez(os.path.join(OUT, "dr_params.npz"), **out)
    import json
    json.dump(CONFIGS, open(os.path.join(OUT, "dr_configs.json"), "w"), indent=1)
    print("wrote", os.path.join(OUT, "dr_params.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
