"""Domain randomisation: BaseTask.apply_randomizations (`tasks/hand_base/base_task.py:229-423`) and the
two noise hooks of BaseTask.step (`:131-132` actions, `:149-150` observations), on the CUDA env.

What is here
  * observations / actions: gaussian or uniform, additive or scaling, linear / constant schedule, white + correlated noise
    (`BT:263-340`) -- parameters computed on the host exactly as the reference does (pinned by tests/golden/dr_params.npz), the
    noise itself by `sdx_dr_randn` / `sdx_dr_noise` (csrc/sdx_dr.cuh) on the env's stream.  The correlated tensor is redrawn
    whenever the parameters are regenerated, as in the reference (a new closure dict without 'corr').
  * sim_params.gravity (`BT:342-355`; the sampling is isaacgym.gymutil.generate_random_samples / apply_random_samples, a
    third-party file that is not under /root/reference -- restated from the published package): three samples are drawn as
    there, the engine has a vertical gravity only and takes the z one (`sdx_set_gravity`).
  * the refresh bookkeeping (`BT:233-249`): everything on the first call, then when `frequency` frames have passed.
What is not: actor_params (per-actor mass / friction / scale / DoF properties through PhysX property setters, `BT:357-409`) --
the contact kernel keeps body parameters per scene, not per env; asking for them raises NotImplementedError.
One deliberate difference: the reference evaluates the refresh inside reset_idx, i.e. on steps in which at least one env
resets; here it is evaluated every step (with thousands of envs some env resets on practically every step, and the fused
step does not tell the host whether one did).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

NONPHYSICAL = ("observations", "actions")


def schedule_scaling(cfg, last_step):                       # BT:269-277
    sched_type = cfg["schedule"] if "schedule" in cfg else None
    sched_step = cfg["schedule_steps"] if "schedule" in cfg else None
    if sched_type == "linear":
        return 1.0 / sched_step * min(last_step, sched_step)
    if sched_type == "constant":
        return 0 if last_step < sched_step else 1
    return 1


def nonphysical_params(cfg, last_step):
    """the four numbers a noise_lambda closes over (BT:279-334) under their reference names, plus the same in the form
    sdx_dr_noise takes.  Additive noise grows from nothing (every number times the schedule factor s); scaling noise grows
    from the identity (spreads times s, centres / bounds interpolated between 1 and their value)."""
    dist, op_type = cfg["distribution"], cfg["operation"]
    if op_type not in ("additive", "scaling") or dist not in ("gaussian", "uniform"):
        raise ValueError(f"unknown operation / distribution {op_type!r} / {dist!r}")
    s = schedule_scaling(cfg, last_step)
    grow = (lambda x: x * s)
    blend = grow if op_type == "additive" else (lambda x: x * s + 1.0 * (1.0 - s))
    first, second = cfg["range"]
    first_c, second_c = cfg.get("range_correlated", [0., 0.])
    if dist == "gaussian":                                  # range = (mean, spread): BT:280-299
        mu, var, mu_corr, var_corr = blend(first), grow(second), blend(first_c), grow(second_c)
        named = {"mu": mu, "var": var, "mu_corr": mu_corr, "var_corr": var_corr}
        a_corr, b_corr, a, b = var_corr, mu_corr, var, mu
    else:                                                   # range = (low, high): BT:303-327
        lo, hi, lo_corr, hi_corr = blend(first), blend(second), blend(first_c), blend(second_c)
        named = {"lo": lo, "hi": hi, "lo_corr": lo_corr, "hi_corr": hi_corr}
        a_corr, b_corr, a, b = hi_corr - lo_corr, lo_corr, hi - lo, lo
    return dict(named, distribution=int(dist == "uniform"), operation=int(op_type == "scaling"), a_corr=a_corr, b_corr=b_corr, a=a, b=b)


def physical_sample(cfg, shape, curr_step, rng):
    """isaacgym.gymutil.generate_random_samples, restated (see module docstring): same schedule treatment as above, numpy draws"""
    dist, op = cfg["distribution"], cfg["operation"]
    s = schedule_scaling(cfg, curr_step)
    grow = (lambda x: x * s)
    blend = grow if op == "additive" else ((lambda x: x * s + 1.0 * (1.0 - s)) if op == "scaling" else (lambda x: x))
    first, second = cfg["range"]
    if dist == "gaussian":
        return rng.normal(blend(first), grow(second) if op in ("additive", "scaling") else second, shape)
    lo, hi = blend(first), blend(second)
    if dist == "loguniform":
        return np.exp(rng.uniform(np.log(lo), np.log(hi), shape))
    if dist == "uniform":
        return rng.uniform(lo, hi, shape)
    raise ValueError(f"unknown distribution {dist!r}")


def check_supported(params):
    """fail loudly on the sections this engine cannot honour (nothing is silently ignored)"""
    for actor, props in (params.get("actor_params") or {}).items():
        live = {k: v for k, v in (props or {}).items() if not (k == "color" and not v)}
        if live:
            raise NotImplementedError(
                f"randomization_params.actor_params.{actor}.{sorted(live)}: per-actor physical properties (BT:357-409) are not "
                "supported -- the contact kernel keeps body parameters per scene, not per env (seqdex_b200/randomization.py)")
    for attr in (params.get("sim_params") or {}):
        if attr != "gravity":
            raise NotImplementedError(f"randomization_params.sim_params.{attr}: only gravity can be randomised (BT:342-355)")
    for k in params:
        if k not in ("frequency", "observations", "actions", "sim_params", "actor_params"):
            raise NotImplementedError(f"randomization_params.{k}: unknown section")


class RefreshSchedule:
    """when apply_randomizations regenerates the non-env parameters (BT:233-249; first_randomization / last_rand_step BT:50-53)"""

    def __init__(self, frequency=1):
        self.frequency, self.first, self.last_rand_step = int(frequency), True, -1

    def due(self, frame):
        do = True if self.first else (frame - self.last_rand_step) >= self.frequency
        if do:
            self.last_rand_step = frame
        self.first = False
        return do


class DomainRandomizer:
    def __init__(self, env, params, seed=22):
        import torch
        check_supported(params)
        self.env, self.params, self.torch = env, params, torch
        self.sched = RefreshSchedule(params.get("frequency", 1))
        self.frame = 0                                       # gym.get_frame_count(sim): simulate() calls so far (BT:240)
        self.seed = (int(seed) * 0x9E3779B97F4A7C15 + 0x44520000) & 0xFFFFFFFFFFFFFFFF
        self.counter = 0                                     # one fresh Philox counter per kernel call
        self.rng = np.random.default_rng(int(seed))          # physical samples: numpy's generator, as in isaacgym.gymutil
        self.state = {}                                      # name -> {"p": parameters, "corr": tensor or None}
        self.gravity0 = float(env.scene.c.gravity_z)
        self.gravity = self.gravity0
        self.randomize_buf = torch.zeros(env.n, dtype=torch.int64, device=env.device)    # BT:67, GS:1642

    def _next(self):
        self.counter = (self.counter + 1) & 0xFFFFFFFF
        return self.counter

    def apply_randomizations(self, reset_buf=None):
        """BT:229-423 (non-env half); returns True when the parameters were regenerated"""
        first = self.sched.first
        do = self.sched.due(self.frame)
        if not first and reset_buf is not None:              # BT:246-249: envs due for physical re-randomisation on their reset
            rand = (self.randomize_buf >= self.sched.frequency) & (reset_buf != 0)
            self.randomize_buf[rand] = 0
        if not do:
            return False
        for name in NONPHYSICAL:
            if name in self.params:
                self.state[name] = {"p": nonphysical_params(dict(self.params[name]), self.frame), "corr": None}
        g = (self.params.get("sim_params") or {}).get("gravity")
        if g is not None:                                    # BT:342-355 -> apply_random_samples(prop, og, 'gravity', ...): x, y, z drawn
            sample = physical_sample(dict(g), 3, self.frame, self.rng)
            self.gravity = self.gravity0 * float(sample[2]) if g["operation"] == "scaling" else self.gravity0 + float(sample[2])
            _lib.check(self.env.L.sdx_set_gravity(self.env.h, ctypes.c_float(self.gravity)))
        return True

    def noise(self, name, src, dst):
        """dr_randomizations[name]['noise_lambda'](src) -> dst (BT:131-132, 149-150); src itself when `name` is not randomised"""
        st = self.state.get(name)
        if st is None:
            return src
        p, L, h = st["p"], self.env.L, self.env.h
        if not src.is_contiguous():
            src = src.contiguous()
        n = src.numel()
        if st["corr"] is None or st["corr"].numel() != n:    # BT:293-296: drawn on the first call after a refresh
            st["corr"] = self.torch.empty(n, dtype=self.torch.float32, device=src.device)
            _lib.check(L.sdx_dr_randn(h, ctypes.c_void_p(st["corr"].data_ptr()), ctypes.c_int64(n), ctypes.c_uint64(self.seed),
                                      ctypes.c_uint32(self._next())))
            st["corr_counter"] = self.counter
        c = self._next()
        _lib.check(L.sdx_dr_noise(h, ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(st["corr"].data_ptr()),
                                  ctypes.c_int64(n), ctypes.c_float(p["a_corr"]), ctypes.c_float(p["b_corr"]), ctypes.c_float(p["a"]),
                                  ctypes.c_float(p["b"]), ctypes.c_int(p["distribution"]), ctypes.c_int(p["operation"]),
                                  ctypes.c_uint64(self.seed), ctypes.c_uint32(c)))
        st["last_counter"] = c
        return dst

    def step_done(self, sim_steps=1):
        self.frame += int(sim_steps)
        self.randomize_buf += 1                              # GS:1642


class RandomizedTaskMixin:
    """what a task class adds around its fused step when cfg['task']['randomize'] is on (GS:106-107, 515-516, 1395-1396; BT:130-150)"""
    randomizer = None

    def _dr_init(self, cfg, seed):
        import torch
        task_cfg = cfg.get("task", {})
        self.randomize = bool(task_cfg.get("randomize", False))
        self.randomization_params = task_cfg.get("randomization_params", {})
        if not self.randomize:
            return
        self.randomizer = DomainRandomizer(self.env, self.randomization_params, seed)
        self.randomizer.apply_randomizations()               # GS:515-516
        self._obs_clean = self.obs_buf                        # the env's own buffer keeps the noise-free frames its history shift needs
        self._obs_noisy = torch.empty_like(self.obs_buf)
        self._act_noisy = torch.empty(self.num_envs, self.num_actions, dtype=torch.float32, device=self.device)

    def _dr_before(self, actions):
        if self.randomizer is None:
            return actions
        self.randomizer.apply_randomizations(self.reset_buf)  # GS:1395-1396 (see module docstring: evaluated every step)
        return self.randomizer.noise("actions", actions.to(self.torch_dtype_f32), self._act_noisy)

    def _dr_after(self):
        if self.randomizer is None:
            return
        self.randomizer.step_done(1 + self.env.last_reset_sim_steps())   # frame count includes the scripted resets' contact steps
        self.obs_buf = self.randomizer.noise("observations", self._obs_clean, self._obs_noisy)

    @property
    def torch_dtype_f32(self):
        import torch
        return torch.float32
