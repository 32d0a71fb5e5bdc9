"""GPU TicTacToe vector environment.

Drop-in for ``SyncVectorEnv([lambda: FlattenMultiDiscreteObservationsWrapper(TicTacToeEnv())] * N,
autoreset_mode=SAME_STEP)`` as built by the reference benchmark (TPB:82-125): the agent plays against a
uniformly random machine (``environments/tiktaktoe_mod.py:96-237``), observations are
``{"observation": base-3 board id (FLT:156-160, UTL:26-29), "action_mask": empty cells}``, rewards +1 / -1 / 0,
finished games restart inside the same step.  Boards are packed 2 bits per cell (bit 18 = agent_mark - 1)
and stepped by one thread per game.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from dist_classicrl_b200 import capi, spaces
from dist_classicrl_b200.environments.custom_env import DeviceVecEnv, _torch


class TicTacToeVecEnv(DeviceVecEnv):
    env_kind = capi.QE_ENV_TTT
    slots = 5

    def __init__(self, num_envs: int, seed: int | None = None, device: int | None = None, output: str = "numpy") -> None:
        super().__init__(num_envs, 3**9, 9, seed=seed, device=device, output=output)
        self.single_action_space = spaces.Discrete(9)
        self.single_observation_space = spaces.Dict(
            {"observation": spaces.Discrete(3**9), "action_mask": spaces.MultiDiscrete([2] * 9)}
        )

    def _reset_kernel(self, u_ptr, slots, seed, t) -> None:
        capi.check(self._lib.qe_ttt_reset(self.env_words.data_ptr(), self.states.data_ptr(), self.mask_bits.data_ptr(), u_ptr,
                                          slots, seed, t, self.agent0, self.num_envs, self._stream()))

    def step(self, actions):
        torch = _torch()
        t = self._rng.next_step()
        u_ptr, slots, seed = self._uniform_args(t)
        n = self.num_envs
        act = self._actions_dev(actions)
        rewards = torch.empty(n, dtype=torch.float32, device=self.device)
        term = torch.empty(n, dtype=torch.uint8, device=self.device)
        capi.check(self._lib.qe_ttt_step(self._err_handle(), self.env_words.data_ptr(), act.data_ptr(), u_ptr, slots, seed, t, self.agent0,
                                         self.states.data_ptr(), self.mask_bits.data_ptr(), rewards.data_ptr(),
                                         term.data_ptr(), n, self._stream()))
        capi.check(self._lib.qe_sync(self._err_handle(), self._stream()))  # raises AssertionError("Invalid move.")
        return self._finish_step(rewards, term)

    def attach(self, algo) -> "TicTacToeVecEnv":
        """Bind to the algorithm whose engine handle carries the device error flags."""
        self._engine = algo
        return self

    def refresh_after_fused(self) -> None:
        b = self.env_words
        occ = (b | (b >> 1)) & 0x15555
        bits = _torch().zeros_like(b)
        for c in range(9):
            bits |= (((~occ) >> (2 * c)) & 1) << c
        self.mask_bits = bits

    # -- board access for tests (mirrors TicTacToeEnv.board / agent_mark) ------------------------------
    @property
    def boards(self) -> np.ndarray:
        w = self.env_words.cpu().numpy().astype(np.int64)
        return np.stack([(w >> (2 * c)) & 3 for c in range(9)], axis=1).astype(np.int8)

    @property
    def agent_marks(self) -> np.ndarray:
        return (((self.env_words.cpu().numpy().astype(np.int64)) >> 18) & 1).astype(np.int8) + 1

    def set_boards(self, boards, agent_marks) -> None:
        torch = _torch()
        boards = np.asarray(boards, dtype=np.int64).reshape(self.num_envs, 9)
        w = np.zeros(self.num_envs, dtype=np.int64)
        for c in range(9):
            w |= boards[:, c] << (2 * c)
        w |= (np.asarray(agent_marks, dtype=np.int64) - 1) << 18
        self.env_words = torch.from_numpy(w.astype(np.int32)).to(self.device)
        radix = 3 ** np.arange(8, -1, -1, dtype=np.int64)
        self.states = torch.from_numpy((boards @ radix).astype(np.int32)).to(self.device)
        self.refresh_after_fused()
