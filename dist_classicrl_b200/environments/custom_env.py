"""Environment contract (reference: ``environments/custom_env.py:15-125``, ``DistClassicRLEnv``).

A multi-agent environment exposes ``num_agents`` (or ``num_envs``), ``reset(seed, options) -> (obs, infos)`` and
``step(actions) -> (obs, rewards float32, terminated bool, truncated bool, infos)`` where ``obs`` is either
``int[N]`` or ``{"observation": int[N], "action_mask": int[N, A]}``.  :class:`DeviceVecEnv` is the base of the
engine's GPU-resident environments; any host environment honouring the contract also works with the runtimes
(through the unfused ``choose_actions`` / ``learn`` path).
"""

from __future__ import annotations

import abc
import ctypes as C
from typing import Any

import numpy as np

from dist_classicrl_b200 import capi
from dist_classicrl_b200.rng import T_INIT, CounterRNG, PredrawnUniforms, is_engine_rng


class DistClassicRLEnv(abc.ABC):
    """Abstract multi-agent environment (same abstract surface as the reference ABC, ENV:31-125)."""

    num_agents: int

    @abc.abstractmethod
    def step(self, actions): ...

    @abc.abstractmethod
    def reset(self, seed: int | None = None, options: dict[str, Any] | None = None): ...

    @abc.abstractmethod
    def close(self) -> None: ...

    @abc.abstractmethod
    def render(self) -> None: ...

    @abc.abstractmethod
    def seed(self, seed: int) -> None: ...

    @abc.abstractmethod
    def get_env_info(self) -> dict[str, Any]: ...

    @abc.abstractmethod
    def get_agent_info(self) -> dict[str, Any]: ...


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("the B200 environments need a CUDA device (there is no CPU fallback)")
    return torch


class _LazyObs(dict):
    def __init__(self, env) -> None:
        super().__init__()
        self._env = env
        self._snap = (env.states.clone(), env.mask_bits.clone())  # the observation of THIS moment (cheap: 8 B / agent)

    def _fill(self) -> None:
        if self._env is not None:
            env, self._env = self._env, None
            super().update(env._obs(*self._snap))
            self._snap = None

    def __getitem__(self, key):
        self._fill()
        return super().__getitem__(key)

    def __iter__(self):
        self._fill()
        return super().__iter__()

    def __len__(self) -> int:
        self._fill()
        return super().__len__()

    def keys(self):
        self._fill()
        return super().keys()

    def items(self):
        self._fill()
        return super().items()

    def values(self):
        self._fill()
        return super().values()

    def __contains__(self, key) -> bool:
        self._fill()
        return super().__contains__(key)

    def get(self, key, default=None):
        self._fill()
        return super().get(key, default)


class DeviceVecEnv(DistClassicRLEnv):
    """Base of the GPU-resident vector environments (state lives in HBM, one thread per agent steps it).

    ``_rng`` is the environment's random stream (slots >= 2 of ``U[t, i, k]``): a
    :class:`~dist_classicrl_b200.rng.CounterRNG` (default) or :class:`~dist_classicrl_b200.rng.PredrawnUniforms`.
    ``step`` takes / returns host NumPy arrays like gymnasium's ``SyncVectorEnv`` (int64 observations and masks)
    unless ``output="torch"`` was requested, in which case everything stays on the device.
    """

    env_kind: int = -1
    slots: int = 2
    dict_obs: bool = True

    def __init__(self, num_envs: int, num_states: int, num_actions: int, seed: int | None = None, device: int | None = None,
                 output: str = "numpy") -> None:
        torch = _torch()
        self._lib = capi.lib()
        self.num_envs = self.num_agents = int(num_envs)
        self.num_states, self.num_actions = int(num_states), int(num_actions)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.output = output
        self._rng = CounterRNG(seed)
        self._resets = 0
        self.agent0 = 0  # global index of this environment's first agent in the uniform stream (replicated multi-GPU mode)
        n = self.num_envs
        self.states = torch.zeros(n, dtype=torch.int32, device=self.device)
        self.states_scratch = torch.zeros(n, dtype=torch.int32, device=self.device)
        self.env_words = torch.zeros(n, dtype=torch.int32, device=self.device)
        self.mask_bits = torch.zeros(n, dtype=torch.int32, device=self.device)
        self.env_seed = 0
        self.term_threshold = 0
        self.episode_len = 0
        self._engine = None  # an OptimalQLearningBase, only used for deferred error reporting

    # -- helpers --------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream().cuda_stream)

    def _uniform_args(self, t: int):
        """(device pointer or None, slots, seed) for stream index ``t``; keeps the upload alive in self._u_dev."""
        rng = self._rng
        if not is_engine_rng(rng):
            raise TypeError("device environments need a CounterRNG or PredrawnUniforms in env._rng")
        if isinstance(rng, PredrawnUniforms):
            torch = _torch()
            row = np.ascontiguousarray(rng.row(t)[: self.num_envs])
            if row.shape[1] < self.slots:
                raise ValueError(f"uniforms need {self.slots} slots")
            self._u_dev = torch.from_numpy(row.view(np.int32)).to(self.device)
            return C.c_void_p(self._u_dev.data_ptr()), row.shape[1], 0
        return None, self.slots, rng.seed

    def _mask_array(self, bits):
        torch = _torch()
        shifts = torch.arange(self.num_actions, device=self.device, dtype=torch.int32)
        return (bits.unsqueeze(1) >> shifts.unsqueeze(0)) & 1

    def _obs_lazy(self):
        """Observation after a fused run: a dict whose entries are built from the device state on first access
        (a caller that hands it straight back to ``run_steps`` never pays for the ``[N, A]`` mask array)."""
        return _LazyObs(self) if self.dict_obs else self._obs()

    def _obs(self, states=None, mask_bits=None):
        states = self.states if states is None else states
        mask_bits = self.mask_bits if mask_bits is None else mask_bits
        if self.output == "torch":
            if not self.dict_obs:
                return states.clone()
            return {"observation": states.clone(), "action_mask": self._mask_array(mask_bits)}
        st = states.cpu().numpy().astype(np.int64)
        if not self.dict_obs:
            return st
        return {"observation": st, "action_mask": self._mask_array(mask_bits).cpu().numpy().astype(np.int64)}

    def _actions_dev(self, actions):
        torch = _torch()
        if isinstance(actions, torch.Tensor):
            return actions.to(device=self.device, dtype=torch.int32).contiguous()
        return torch.from_numpy(np.ascontiguousarray(actions, dtype=np.int32)).to(self.device)

    def _finish_step(self, rewards, term):
        n = self.num_envs
        if self.output == "torch":
            torch = _torch()
            return self._obs(), rewards, term.bool(), torch.zeros(n, dtype=torch.bool, device=self.device), {}
        return self._obs(), rewards.cpu().numpy(), term.cpu().numpy().astype(bool), np.zeros(n, dtype=bool), {}

    # -- gym-like surface -----------------------------------------------------------------------------
    def seed(self, seed: int) -> None:
        self._rng = CounterRNG(seed)
        self._resets = 0

    def reset(self, seed: int | None = None, options: dict[str, Any] | None = None):
        if seed is not None:
            self.seed(seed)
        if isinstance(self._rng, PredrawnUniforms):
            raise TypeError("reset() of a device environment draws from a CounterRNG; use reset_with(uniforms) instead")
        t = (T_INIT - self._resets) & 0xFFFFFFFF
        self._resets += 1
        self._reset_kernel(None, self.slots, self._rng.seed, t)
        return self._obs(), {}

    def reset_with(self, uniforms: np.ndarray):
        """Reset from a caller-supplied ``uint32[N, slots]`` row of uniforms (parity tests)."""
        torch = _torch()
        row = np.ascontiguousarray(uniforms, dtype=np.uint32)
        d = torch.from_numpy(row.view(np.int32)).to(self.device)
        self._reset_kernel(C.c_void_p(d.data_ptr()), row.shape[1], 0, 0)
        torch.cuda.current_stream().synchronize()
        return self._obs(), {}

    def close(self) -> None:
        return None

    def render(self) -> None:
        return None

    def get_env_info(self) -> dict[str, Any]:
        return {"num_states": self.num_states, "num_actions": self.num_actions, "num_agents": self.num_agents}

    def get_agent_info(self) -> dict[str, Any]:
        return {"num_agents": self.num_agents}

    def agents_struct(self, episode_returns) -> capi.QeAgents:
        """``qe_agents_t`` view of this environment for the fused loop."""
        words = None if self.env_kind == capi.QE_ENV_MDP else self.env_words.data_ptr()
        return capi.QeAgents(self.env_kind, self.num_envs, self.states.data_ptr(), self.states_scratch.data_ptr(), words,
                             episode_returns.data_ptr(), self.env_seed, self.episode_len, self.term_threshold)

    def _err_handle(self):
        """Engine handle whose device error flag the step kernels report into (a private 1x1 engine unless the
        environment was attached to an algorithm)."""
        if self._engine is None:
            from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase

            self._engine = OptimalQLearningBase(1, 1, device=self.device.index)
        return self._engine.handle

    def attach(self, algo):
        self._engine = algo
        return self

    def refresh_after_fused(self) -> None:
        """Recompute derived device state (action masks) after the fused loop advanced ``states``/``env_words``."""

    def _reset_kernel(self, u_ptr, slots, seed, t) -> None:
        raise NotImplementedError
