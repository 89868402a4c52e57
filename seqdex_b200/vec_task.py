"""VecEnv adapter with the reference's surface (tasks/hand_base/vec_task_rlgames.py, VR:18-213): what an
rl_games-style algorithm calls.  Boundary B1 of SURVEY.md section 8b."""
from __future__ import annotations

import numpy as np
import torch


class Box:
    """the two attributes of gym.spaces.Box rl_games reads (gym is not a dependency here)"""

    def __init__(self, low, high):
        self.low, self.high = np.asarray(low, np.float32), np.asarray(high, np.float32)
        self.shape = self.low.shape
        self.dtype = np.float32


class VecTask:
    def __init__(self, task, rl_device, clip_observations=5.0, clip_actions=1.0):
        self.task = task
        self.num_environments = task.num_envs
        self.num_agents = 1
        self.num_observations = task.num_obs
        self.num_states = task.num_states
        self.num_actions = task.num_actions
        self.obs_space = Box(np.ones(self.num_obs) * -np.inf, np.ones(self.num_obs) * np.inf)
        self.state_space = Box(np.ones(self.num_states) * -np.inf, np.ones(self.num_states) * np.inf)
        self.act_space = Box(np.ones(self.num_actions) * -1., np.ones(self.num_actions) * 1.)
        self.clip_obs = clip_observations
        self.clip_actions = clip_actions
        self.rl_device = task.device          # VR:33 (ignores the --rl_device flag, like the reference)
        self.info = {"action_space": self.act_space, "observation_space": self.obs_space,
                     "state_space": self.state_space, "agents": 1}

    def has_action_masks(self):
        return False

    def seed(self, seed):
        pass

    def set_train_info(self, env_frames, *args, **kwargs):
        pass

    def get_env_state(self):
        return None

    def set_env_state(self, env_state):
        pass

    @property
    def get_number_of_agents(self):
        return self.num_agents

    @property
    def observation_space(self):
        return self.obs_space

    @property
    def action_space(self):
        return self.act_space

    @property
    def num_envs(self):
        return self.num_environments

    @property
    def num_acts(self):
        return self.num_actions

    @property
    def num_obs(self):
        return self.num_observations

    def get_env_info(self):
        return self.info


class RLgamesVecTaskPython(VecTask):
    def get_state(self):
        return torch.clamp(self.task.states_buf, -self.clip_obs, self.clip_obs)

    def step(self, actions):                                                         # VR:165-177
        actions_tensor = torch.clamp(actions, -self.clip_actions, self.clip_actions)
        self.task.step(actions_tensor)
        obs_dict = {"obs": torch.clamp(self.task.obs_buf, -self.clip_obs, self.clip_obs),
                    "states": torch.clamp(self.task.states_buf, -self.clip_obs, self.clip_obs)}
        return obs_dict, self.task.rew_buf, self.task.reset_buf, self.task.extras

    def step_into(self, actions, obs_out, states_out):
        """step() writing the clamped observations straight into caller buffers (same values as step(); saves the two
        temporaries torch.clamp allocates every step).  Actions are clamped inside the pre-physics kernel."""
        if getattr(self.task, "randomizer", None) is not None:    # domain randomisation: noise goes on the CLAMPED actions (VR:166, BT:131)
            self.task.step(torch.clamp(actions, -self.clip_actions, self.clip_actions))
            torch.clamp(self.task.obs_buf, -self.clip_obs, self.clip_obs, out=obs_out)
            self.task.env.clamped_copy("STATES", states_out, self.clip_obs)
            return self.task.rew_buf, self.task.reset_buf, self.task.extras
        self.task.step(actions)
        self.task.env.clamped_copy("OBS", obs_out, self.clip_obs)
        self.task.env.clamped_copy("STATES", states_out, self.clip_obs)
        return self.task.rew_buf, self.task.reset_buf, self.task.extras

    def reset(self):                                                                 # VR:179-192
        actions = 0.01 * (1 - 2 * torch.rand([self.task.num_envs, self.task.num_actions], dtype=torch.float32, device=self.rl_device))
        self.task.step(actions)
        return {"obs": torch.clamp(self.task.obs_buf, -self.clip_obs, self.clip_obs),
                "states": torch.clamp(self.task.states_buf, -self.clip_obs, self.clip_obs)}
