"""List-of-environments vector wrapper (drop-in for ``wrappers/dummy_vec_wrapper.py:9-101``): steps host environments
one by one and stacks their results; no autoreset (the wrapped environments decide what a finished episode does)."""

from __future__ import annotations

import numpy as np


class DummyVecWrapper:
    def __init__(self, envs: list) -> None:
        self.env = envs[0]
        self.envs = envs
        self.num_envs = len(envs)
        self.observation_space = envs[0].observation_space
        self.action_space = envs[0].action_space

    def __getattr__(self, name):  # gymnasium.Wrapper forwards unknown attributes to the wrapped environment
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, seed: int | None = None, options: dict | None = None):
        obs, infos = [], []
        for env in self.envs:
            o, info = env.reset(seed=seed, options=options)
            obs.append(o)
            infos.append(info)
        return np.array(obs), infos

    def step(self, actions):
        if len(actions) != len(self.envs):
            raise ValueError("zip() arguments have different lengths")  # zip(strict=True), dummy_vec_wrapper.py:76
        cols: tuple[list, ...] = ([], [], [], [], [])
        for env, action in zip(self.envs, actions):
            for col, x in zip(cols, env.step(action)):
                col.append(x)
        return np.array(cols[0]), np.array(cols[1]), np.array(cols[2]), np.array(cols[3]), cols[4]

    def close(self) -> None:
        for env in self.envs:
            env.close()

    def render(self) -> None:
        for env in self.envs:
            env.render()
