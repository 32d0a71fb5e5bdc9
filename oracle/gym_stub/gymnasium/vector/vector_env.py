"""SAME_STEP ``SyncVectorEnv`` restatement (last section)."""
from __future__ import annotations

import enum

import numpy as np

from gymnasium import spaces


class AutoresetMode(enum.Enum):
    NEXT_STEP = "NextStep"
    SAME_STEP = "SameStep"
    DISABLED = "Disabled"


class VectorEnv:
    num_envs: int = 0


def _stack(space, items):
    if isinstance(space, spaces.Dict):
        return {k: _stack(sub, [it[k] for it in items]) for k, sub in space.spaces.items()}
    return np.asarray(items, dtype=np.int64)  # Discrete / MultiDiscrete -> int64


class SyncVectorEnv(VectorEnv):
    def __init__(self, env_fns, autoreset_mode=AutoresetMode.NEXT_STEP):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        self.autoreset_mode = autoreset_mode
        assert autoreset_mode == AutoresetMode.SAME_STEP, "stub implements SAME_STEP only"
        self.single_observation_space = self.envs[0].observation_space
        self.single_action_space = self.envs[0].action_space

    def reset(self, *, seed=None, options=None):
        obs = []
        for i, env in enumerate(self.envs):
            o, _ = env.reset(seed=None if seed is None else seed + i, options=options)
            obs.append(o)
        return _stack(self.single_observation_space, obs), {}

    def step(self, actions):
        obs, rewards, terms, truncs = [], [], [], []
        for env, action in zip(self.envs, actions):
            o, r, term, trunc, _ = env.step(action)
            if term or trunc:  # SAME_STEP: reset inside the step, return the reset observation
                o, _ = env.reset()
            obs.append(o)
            rewards.append(r)
            terms.append(term)
            truncs.append(trunc)
        return (
            _stack(self.single_observation_space, obs),
            np.asarray(rewards, dtype=np.float64),
            np.asarray(terms, dtype=bool),
            np.asarray(truncs, dtype=bool),
            {},
        )

    def close(self):
        for env in self.envs:
            env.close()
