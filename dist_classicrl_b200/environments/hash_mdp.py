"""Synthetic integer-hash tabular MDP on the GPU (BASELINE configs 3-5; new -- SURVEY 8d).

``h = mix32(s*A + a + seedmix, 0)``, ``s' = (h*S) >> 32``, ``r = float32((mix32(h,1) >> 8) * 2**-24) * 2 - 1``,
``terminated = U[t,i,2] < ceil(p_term * 2**32)``, restart state ``(U[t,i,3]*S) >> 32`` (SAME_STEP),
legal actions of a state = ``(mix32(s + seedmix, 2) & (2**A - 1)) | 1``.  Requires ``S*A < 2**32``, ``A <= 32``.
"""

from __future__ import annotations

import math

from dist_classicrl_b200 import capi, spaces
from dist_classicrl_b200.environments.custom_env import DeviceVecEnv, _torch


class HashMDPVecEnv(DeviceVecEnv):
    env_kind = capi.QE_ENV_MDP
    slots = 4

    def __init__(self, num_envs: int, num_states: int, num_actions: int, env_seed: int = 0, p_term: float = 0.05,
                 seed: int | None = None, device: int | None = None, output: str = "numpy") -> None:
        if num_states * num_actions >= 2**32 or num_actions > 32:
            raise ValueError("the hash MDP needs S*A < 2**32 and A <= 32")
        super().__init__(num_envs, num_states, num_actions, seed=env_seed if seed is None else seed, device=device, output=output)
        self.env_seed = int(env_seed)
        self.p_term = p_term
        self.term_threshold = int(math.ceil(p_term * 2.0**32))
        self.single_action_space = spaces.Discrete(num_actions)
        self.single_observation_space = spaces.Dict(
            {"observation": spaces.Discrete(num_states), "action_mask": spaces.MultiDiscrete([2] * num_actions)}
        )

    def attach(self, algo) -> "HashMDPVecEnv":
        self._engine = algo
        return self

    def _reset_kernel(self, u_ptr, slots, seed, t) -> None:
        capi.check(self._lib.qe_mdp_reset(self.states.data_ptr(), self.mask_bits.data_ptr(), self.num_states, self.num_actions,
                                          self.env_seed, u_ptr, slots, seed, t, self.agent0, self.num_envs, self._stream()))

    def step(self, actions):
        torch = _torch()
        t = self._rng.next_step()
        u_ptr, slots, seed = self._uniform_args(t)
        n = self.num_envs
        act = self._actions_dev(actions)
        rewards = torch.empty(n, dtype=torch.float32, device=self.device)
        term = torch.empty(n, dtype=torch.uint8, device=self.device)
        capi.check(self._lib.qe_mdp_step(self._err_handle(), self.states.data_ptr(), act.data_ptr(), self.num_states,
                                         self.num_actions, self.env_seed, self.term_threshold, u_ptr, slots, seed, t, self.agent0,
                                         self.mask_bits.data_ptr(), rewards.data_ptr(), term.data_ptr(), n, self._stream()))
        capi.check(self._lib.qe_sync(self._err_handle(), self._stream()))
        return self._finish_step(rewards, term)

    def refresh_after_fused(self) -> None:
        # masks are a pure function of the state
        capi.check(self._lib.qe_mdp_masks(self.states.data_ptr(), self.mask_bits.data_ptr(), self.num_actions, self.env_seed,
                                          self.num_envs, self._stream()))
