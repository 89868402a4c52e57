"""Policy sequencing on top of the CUDA PPO engine.

  * ``NNController``            -- utils/robot_controller/nn_controller.py:7-58: a frozen rl_games actor used as an inner
                                   policy by the chain tasks (``insert_policy.predict(obs)``, tool_positioning_chain TC:1735-1768);
  * ``LegoVecTaskPython``       -- tasks/hand_base/vec_task_lego.py:190-237: VecEnv whose ``step`` hands back the per-phase
                                   observations of a sequenced task through ``extras['before_*' / 'after_*']``;
  * ``PolicySequencingRunner``  -- policy_sequencing/policy_seq_runner.py PSR:39-373: two PPO agents on ONE env; the agent
                                   whose phase is active (``progress_buf[0] < before_episode_length``) acts and is trained.

Only launch sequencing lives here; every contraction / elementwise op is a kernel of csrc/sdx_ppo.cu.
"""
from __future__ import annotations

import ctypes
import time

import torch

from . import _lib, checkpoint
from .ppo import MLP, A2CAgent, PPOConfig, _p, _stream
from .vec_task import Box, VecTask


class NNController:
    """rl_games actor (continuous_a2c_logstd, ``separate: True``, fixed sigma) restored from a ``.pth`` and evaluated with
    the tensor-core MLP.  ``units`` are the yaml's ``network.mlp.units`` (robot_controller/network.yaml: [512, 256, 128];
    grasp_network.yaml / insert_network.yaml: [1024, 512, 256])."""

    def __init__(self, num_actors=1, units=(512, 256, 128), obs_dim=81, actions_num=23, device=0, seed=0):
        self.one_step_obs_dim, self.actions_num, self.units = obs_dim, actions_num, tuple(units)
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        self.L = _lib.load()
        self.rows = (max(int(num_actors), 128) + 7) // 8 * 8
        self.model = MLP(obs_dim, actions_num, self.rows, has_sigma=True, device=self.device.index or 0, seed=seed, hidden=self.units)
        self.states = None            # rnn states: the reference's networks have none
        self.time_step = 0
        self.seed, self.counter = seed, 0
        self.last_action = None
        self._act = torch.zeros(self.rows, actions_num, device=self.device)
        self._nlp = torch.zeros(self.rows, device=self.device)
        self._obs = torch.zeros(self.rows, obs_dim, device=self.device)

    def load(self, fn):
        ck = checkpoint.load_checkpoint(fn)
        flat, _ = checkpoint.actor_flat(ck["model"], self.one_step_obs_dim, self.actions_num, self.units)
        self.model.load_flat(flat)

    def predict(self, observation, deterministic=False):
        """nn_controller.py:27-53: mu (deterministic) or mu + sigma*eps, clipped to [-1, 1]"""
        if not isinstance(observation, torch.Tensor):
            observation = torch.as_tensor(observation)
        obs = observation.to(self.device)
        obs = obs.float() / 255.0 if obs.dtype == torch.uint8 else obs.float()      # _preproc_obs, nn_controller.py:15-25
        if obs.dim() == 1:
            obs = obs.unsqueeze(0)
        M = obs.shape[0]
        assert M <= self.rows, "NNController was sized for fewer actors"
        Mp = (M + 7) // 8 * 8                                                        # rows are staged 8 at a time
        self._obs[:M].copy_(obs)
        mu = self.model.forward(self._obs[:Mp])[:M]
        if deterministic:
            cur = mu.clone()
        else:
            logstd = self.model.params[self.model.nparams - self.actions_num:]
            _lib.check(self.L.sdx_ppo_sample(_p(self.model.out), _p(logstd), M, self.actions_num, ctypes.c_uint64(self.seed), self.counter,
                                             _p(self._act), _p(self._nlp), _stream()))
            self.counter += 1
            cur = self._act[:M].clone()
        self.last_action = cur
        self.time_step += 1
        return torch.clip(cur, -1.0, 1.0).detach()


class LegoVecTaskPython(VecTask):
    """vec_task_lego.py:190-237.  The task must publish ``extras['before_obs', 'before_states', 'after_obs', 'after_states',
    'before_rew_buf', 'after_rew_buf', 'before_reset_buf', 'after_reset_buf']`` and the per-phase sizes
    ``grasping_num_obs / grasping_num_states / insertion_num_obs / insertion_num_states`` (vec_task_lego.py:49-59)."""

    def __init__(self, task, rl_device, clip_observations=5.0, clip_actions=1.0):
        super().__init__(task, rl_device, clip_observations, clip_actions)
        import numpy as np
        mk = lambda n: Box(np.ones(n) * -np.inf, np.ones(n) * np.inf)
        self.grasp_info = {"action_space": self.act_space, "observation_space": mk(task.grasping_num_obs),
                           "state_space": mk(task.grasping_num_states), "agents": 1}
        self.insert_info = {"action_space": self.act_space, "observation_space": mk(task.insertion_num_obs),
                            "state_space": mk(task.insertion_num_states), "agents": 1}

    def get_grasp_env_info(self):
        return self.grasp_info

    def get_insert_env_info(self):
        return self.insert_info

    def _clamp_extras(self):
        ex = self.task.extras
        for k in ("before_obs", "before_states", "after_obs", "after_states"):
            ex[k] = torch.clamp(ex[k], -self.clip_obs, self.clip_obs)
        return ex

    def step(self, actions):
        self.task.step(torch.clamp(actions, -self.clip_actions, self.clip_actions))
        return {}, self.task.rew_buf, self.task.reset_buf, self._clamp_extras()

    def reset(self):
        actions = 0.01 * (1 - 2 * torch.rand([self.task.num_envs, self.task.num_actions], dtype=torch.float32, device=self.rl_device))
        self.task.step(actions)
        ex = self._clamp_extras()
        return {"obs": ex["after_obs"], "states": ex["after_states"], "before_obs": ex["before_obs"], "before_states": ex["before_states"]}


class PolicySequencingRunner:
    """PSR:39-373 with the CUDA agents.  ``env`` is a ``LegoVecTaskPython``; agent 0 ("before") owns the steps with
    ``progress_buf[0] < before_episode_length`` (PSR:61, 222, 196), agent 1 ("after") the rest."""

    def __init__(self, env, cfg_before: PPOConfig | None = None, cfg_after: PPOConfig | None = None, device=0,
                 before_checkpoint="", after_checkpoint="", before_episode_length=100, dist_group=None):
        self.env = env
        self.before_episode_length = before_episode_length
        views = [_PhaseView(env, env.get_grasp_env_info()), _PhaseView(env, env.get_insert_env_info())]
        self.agents = [A2CAgent(views[0], cfg_before, device, dist_group), A2CAgent(views[1], cfg_after, device, dist_group)]
        for agent, ck in zip(self.agents, (before_checkpoint, after_checkpoint)):   # PSR:74-75, 86-87 (_restore)
            if ck:
                agent.restore(ck)
            agent.last_mean_rewards = -100500
        first = env.reset()                                                           # PSR:100-106
        self.agents[1].set_obs(first["obs"], first["states"])
        self.agents[0].set_obs(first["before_obs"], first["before_states"])
        self.total_time = 0.0

    def _phase(self):
        """PSR:196, 222: ONE host read of progress_buf[0] decides the phase for every env (the chain runs in lockstep)"""
        return 0 if int(self.env.task.progress_buf[0]) < self.before_episode_length else 1

    def rl_games_play_steps(self):
        step_time = 0.0
        H = self.agents[0].H
        for n in range(H):
            k = self._phase()
            agent, pre = self.agents[k], ("before", "after")[k]
            actions = agent.act(n)
            t0 = time.time()
            _, _, _, infos = self.env.step(actions)
            agent.set_obs(infos[pre + "_obs"], infos[pre + "_states"])               # PSR:235-236, 256-257
            step_time += time.time() - t0
            agent.record(n, infos[pre + "_rew_buf"], infos[pre + "_reset_buf"].float())   # PSR:237-238, 258-259
        for agent in self.agents:                                                      # PSR:267-268: both batches are closed
            agent.finish_rollout()
        return step_time

    def rl_games_train_epoch(self):
        t0 = time.time()
        step_time = self.rl_games_play_steps()
        t1 = time.time()
        k = self._phase()                                                              # PSR:196-201: only the active agent learns
        info = self.agents[k].update()
        t2 = time.time()
        return step_time, t1 - t0, t2 - t1, t2 - t0, info, ("before", "after")[k]

    def run(self, max_epochs=1):
        out = []
        for _ in range(max_epochs):
            step_time, play_time, update_time, sum_time, info, name = self.rl_games_train_epoch()
            self.total_time += sum_time
            frames = self.agents[0].B
            info = dict(info, trained=name, fps_step=frames / max(step_time, 1e-6), fps_step_inference=frames / play_time,
                        fps_total=frames / sum_time)                                    # PSR:133-139
            out.append(info)
        return out


class _PhaseView:
    """what an A2CAgent needs to know of its env: sizes of ITS phase's observation / state (PSR:41-55 env_info override)"""

    def __init__(self, env, info):
        self.env = env
        self.num_envs = env.num_envs
        self.num_actions = env.num_actions
        self.num_obs = int(info["observation_space"].shape[0])
        self.num_states = int(info["state_space"].shape[0])

    def get_env_state(self):
        return self.env.get_env_state()
