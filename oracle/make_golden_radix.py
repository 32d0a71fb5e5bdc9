"""Generate ``tests/golden/radix.npz`` from the REAL reference's ``utils.py`` and flatten wrappers (TEST INFRASTRUCTURE;
build container only: ``python oracle/make_golden_radix.py``).  gymnasium is absent here, so the reference modules are
imported on top of ``oracle/gym_stub`` like ``oracle/make_golden.py`` does."""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "gym_stub"))
sys.path.insert(0, "/root/reference/src")

import gymnasium  # noqa: E402
from dist_classicrl import utils as ref  # noqa: E402
from dist_classicrl.wrappers.flatten_multidiscrete_wrapper import (  # noqa: E402
    FlattenMultiDiscreteActionsWrapper, FlattenMultiDiscreteObservationsWrapper)
from gymnasium import spaces  # noqa: E402


class GridEnv(gymnasium.Env):
    """MultiDiscrete observations AND actions: the observation is the previous action plus a step counter digit."""

    def __init__(self, obs_nvec, act_nvec, dict_obs):
        self.obs_nvec, self.act_nvec, self.dict_obs = np.asarray(obs_nvec), np.asarray(act_nvec), dict_obs
        sub = spaces.MultiDiscrete(self.obs_nvec)
        self.observation_space = spaces.Dict({"observation": sub, "action_mask": spaces.MultiDiscrete([2] * 3)}) if dict_obs else sub
        self.action_space = spaces.MultiDiscrete(self.act_nvec)
        self.t = 0
        self.seen = []

    def _obs(self, vec):
        vec = np.asarray(vec, dtype=np.int32) % self.obs_nvec
        return {"observation": vec, "action_mask": np.ones(3, dtype=np.int8)} if self.dict_obs else vec

    def reset(self, seed=None, options=None):
        self.t = 0
        return self._obs(np.arange(len(self.obs_nvec))), {}

    def step(self, action):
        self.seen.append(np.asarray(action).copy())
        self.t += 1
        vec = np.resize(np.asarray(action), len(self.obs_nvec)) + self.t
        return self._obs(vec), float(self.t), False, False, {}


def main():
    out = {}
    rng = np.random.default_rng(0)
    for name, nvec in (("ttt", [3] * 9), ("mixed", [4, 1, 7, 2, 5]), ("one", [11]), ("wide", [2] * 30)):
        nvec = np.asarray(nvec, dtype=np.int32)
        radix = ref.compute_radix(nvec)
        n = 257
        vecs = (rng.integers(0, 1 << 30, size=(n, len(nvec))) % nvec).astype(np.int32)
        codes = ref.encode_multi_discretes(vecs, radix)
        singles = np.array([ref.encode_multi_discrete(v, radix) for v in vecs[:16]])
        back = ref.decode_to_multi_discretes(nvec, codes.reshape(-1, 1), radix)
        back1 = np.stack([ref.decode_to_multi_discrete(nvec, int(c), radix) for c in codes[:16]])
        out.update({f"{name}__nvec": nvec, f"{name}__radix": radix, f"{name}__vectors": vecs, f"{name}__codes": codes,
                    f"{name}__codes_single": singles, f"{name}__decoded": back, f"{name}__decoded_single": back1})
    # the wrappers, stacked like the benchmark stacks them (TPB:109-123): flat actions in, flat observations out
    for name, dict_obs in (("wrap_plain", False), ("wrap_dict", True)):
        env = GridEnv([3, 4, 5, 2], [2, 3, 4], dict_obs)
        wrapped = FlattenMultiDiscreteObservationsWrapper(FlattenMultiDiscreteActionsWrapper(env))
        obs, _ = wrapped.reset()
        flat_obs = [obs["observation"] if dict_obs else obs]
        actions = rng.integers(0, 24, size=12)
        for a in actions:
            obs, *_ = wrapped.step(int(a))
            flat_obs.append(obs["observation"] if dict_obs else obs)
        out.update({f"{name}__actions": actions, f"{name}__flat_obs": np.asarray(flat_obs), f"{name}__inner_actions": np.stack(env.seen),
                    f"{name}__n_obs": np.asarray(wrapped.observation_space.spaces["observation"].n if dict_obs else wrapped.observation_space.n),
                    f"{name}__n_act": np.asarray(wrapped.action_space.n)})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "radix.npz"), **out)
    print("wrote radix.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
