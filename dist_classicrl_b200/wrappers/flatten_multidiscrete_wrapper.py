"""Flatten wrappers (drop-in for ``wrappers/flatten_multidiscrete_wrapper.py:21-161``): a MultiDiscrete action or
observation space becomes one Discrete space through mixed-radix coding (``utils.compute_radix``).

Like the reference they wrap a single environment (one vector per call: host integer arithmetic); unlike it they also
wrap *vector* environments whose observations are ``[n, dims]`` arrays or CUDA tensors and whose actions are ``[n]``
indices -- those go through the batched GPU kernels (``utils.encode_multi_discretes`` / ``decode_to_multi_discretes``).
"""

from __future__ import annotations

import numpy as np

from dist_classicrl_b200 import spaces
from dist_classicrl_b200.utils import (compute_radix, decode_to_multi_discrete, decode_to_multi_discretes, encode_multi_discrete,
                                       encode_multi_discretes)


def _is_space(obj, name: str) -> bool:
    return type(obj).__name__ == name  # ours or gymnasium's


class _Wrapper:
    def __init__(self, env) -> None:
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)


class FlattenMultiDiscreteActionsWrapper(_Wrapper):
    """Discrete(prod(nvec)) actions for an environment with MultiDiscrete(nvec) actions (FLT:21-76)."""

    def __init__(self, env) -> None:
        super().__init__(env)
        action_space = env.action_space
        assert _is_space(action_space, "MultiDiscrete") or _is_space(action_space, "Discrete"), (
            f"Expected MultiDiscrete or Discrete action space, got {type(env.action_space)}.")
        assert _is_space(action_space, "MultiDiscrete"), "Expected MultiDiscrete action space."
        self.action_radix = compute_radix(action_space.nvec)
        self.action_nvec = action_space.nvec
        self.action_space = spaces.Discrete(np.prod(action_space.nvec))

    def action(self, action):
        """A flat action (or ``[n]`` flat actions of a vector environment) -> MultiDiscrete vector(s)."""
        if np.ndim(action) == 0 and not hasattr(action, "is_cuda"):
            return decode_to_multi_discrete(self.action_nvec, action, self.action_radix)
        return decode_to_multi_discretes(self.action_nvec, action, self.action_radix)

    def step(self, action):
        return self.env.step(self.action(action))


class FlattenMultiDiscreteObservationsWrapper(_Wrapper):
    """Discrete(prod(nvec)) observations for MultiDiscrete(nvec) observations, bare or under the ``"observation"`` key
    of a Dict space (FLT:78-161).  Like the reference it rewrites the wrapped environment's Dict space and the
    observation dictionaries in place (FLT:120-125, 157-160)."""

    def __init__(self, env) -> None:
        super().__init__(env)
        observation_space = env.observation_space
        if _is_space(observation_space, "Dict"):
            assert "observation" in observation_space.spaces, "Expected 'observation' key in observation space."
            sub = observation_space.spaces["observation"]
            assert _is_space(sub, "MultiDiscrete"), "Expected MultiDiscrete observation space."
            self.observation_radix = compute_radix(sub.nvec)
            self.observation_nvec = sub.nvec
            self.observation_space = observation_space
            self.observation_space.spaces["observation"] = spaces.Discrete(np.prod(sub.nvec))
        else:
            assert _is_space(observation_space, "MultiDiscrete") or _is_space(observation_space, "Discrete"), (
                f"Expected MultiDiscrete or Discrete observation space, got {type(env.observation_space)}.")
            assert _is_space(observation_space, "MultiDiscrete"), "Expected MultiDiscrete observation space."
            self.observation_radix = compute_radix(observation_space.nvec)
            self.observation_nvec = observation_space.nvec
            self.observation_space = spaces.Discrete(np.prod(observation_space.nvec))

    def _encode(self, vec):
        if np.ndim(vec) == 2 or (hasattr(vec, "is_cuda") and vec.ndim == 2):
            return encode_multi_discretes(vec, self.observation_radix)
        return encode_multi_discrete(vec, self.observation_radix)

    def observation(self, observation):
        if isinstance(observation, dict):
            observation["observation"] = self._encode(observation["observation"])
            return observation
        return self._encode(observation)

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return self.observation(obs), reward, terminated, truncated, info
