"""Rigged two-armed bandit, vectorised (reference: ``environments/rigged_two_armed_bandit.py:6-94`` inside
``DummyVecWrapper``, ``wrappers/dummy_vec_wrapper.py:58-91``): one state, reward = action, the episode ends every
``episode_len`` steps.  Observations are plain ``int64[N]`` arrays (no action masks)."""

from __future__ import annotations

import numpy as np

from dist_classicrl_b200 import capi, spaces
from dist_classicrl_b200.environments.custom_env import DeviceVecEnv, _torch


class RiggedTwoArmedBanditVecEnv(DeviceVecEnv):
    env_kind = capi.QE_ENV_BANDIT
    slots = 2
    dict_obs = False

    def __init__(self, num_envs: int, episode_len: int = 10, device: int | None = None, output: str = "numpy") -> None:
        super().__init__(num_envs, 1, 2, seed=0, device=device, output=output)
        self.episode_len = int(episode_len)
        self.single_action_space = spaces.Discrete(2)
        self.single_observation_space = spaces.Discrete(1)

    def attach(self, algo) -> "RiggedTwoArmedBanditVecEnv":
        self._engine = algo
        return self

    def _reset_kernel(self, u_ptr, slots, seed, t) -> None:
        self.env_words.zero_()
        self.states.zero_()

    def reset(self, seed=None, options=None):
        self._reset_kernel(None, 0, 0, 0)
        return self._obs(), [{} for _ in range(self.num_envs)]

    def step(self, actions):
        torch = _torch()
        act = self._actions_dev(actions)
        bad = ((act < 0) | (act > 1)).any()
        if bool(bad):
            raise AssertionError(f"Invalid action: {actions}")
        self.env_words += 1
        term = self.env_words >= self.episode_len
        self.env_words = torch.where(term, torch.zeros_like(self.env_words), self.env_words)
        rewards = act.to(torch.float32)
        out = self._finish_step(rewards, term.to(torch.uint8))
        return (*out[:4], [{} for _ in range(self.num_envs)])


def make_bandit_vec_env(n_envs: int, episode_len: int = 10, **kw) -> RiggedTwoArmedBanditVecEnv:
    """Counterpart of ``utils._make_dummy_vec_env(n, RiggedTwoArmedBanditEnv, {...})`` (UTL:118-139)."""
    return RiggedTwoArmedBanditVecEnv(n_envs, episode_len=episode_len, **kw)


__all__ = ["RiggedTwoArmedBanditVecEnv", "make_bandit_vec_env", "np"]
