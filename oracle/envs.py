"""NumPy restatement of the bundled vector environments (TEST INFRASTRUCTURE).

* :class:`TicTacToeVec` -- ``TicTacToeEnv`` (TTT:67-237) wrapped by
  ``FlattenMultiDiscreteObservationsWrapper`` (FLT:139-161, radix UTL:12-48) inside
  gymnasium ``SyncVectorEnv(autoreset_mode=SAME_STEP)`` (third party, restated in
  SURVEY Appendix B; "parity unpinned" -- no reference test covers it).
* :class:`BanditVec`    -- ``RiggedTwoArmedBanditEnv`` (rigged_two_armed_bandit.py:55-80)
  in ``DummyVecWrapper`` (dummy_vec_wrapper.py:58-91; no autoreset).
* :class:`HashMDPVec`   -- the synthetic integer-hash tabular MDP of BASELINE configs 3-5
  (new, no reference counterpart).

Every env consumes slots 2.. of the pre-drawn stream ``U[t]`` (``oracle.rng``).
``step`` returns ``(obs, rewards float32, terminated bool, truncated bool, infos)``
with ``obs = {"observation": int64[N], "action_mask": int64[N, A]}`` (dict envs)
or ``int64[N]``.
"""

from __future__ import annotations

import numpy as np

from oracle.rng import SLOT_ENV0, SLOT_ENV1, SLOT_ENV2, draw_uniforms, mix32, pick

T_INIT = 0xFFFFFFFF  # stream index used by the initial reset

_LINES = np.array(
    [[0, 1, 2], [3, 4, 5], [6, 7, 8], [0, 3, 6], [1, 4, 7], [2, 5, 8], [0, 4, 8], [2, 4, 6]]
)  # TTT:216-237
_RADIX = 3 ** np.arange(8, -1, -1, dtype=np.int64)  # compute_radix([3]*9), UTL:26-29


def ttt_encode(boards: np.ndarray) -> np.ndarray:
    """``state = dot(board, radix)``, cell 0 most significant (FLT:156-160, UTL:48)."""
    return boards.astype(np.int64) @ _RADIX


def ttt_has_line(boards: np.ndarray, mark: np.ndarray) -> np.ndarray:
    """``_check_winner() == mark`` per env (TTT:132, 216-237)."""
    cells = boards[:, _LINES]  # [N, 8, 3]
    return (cells == mark[:, None, None]).all(axis=2).any(axis=1)


class TicTacToeVec:
    num_actions = 9
    num_states = 3**9
    slots = 5

    def __init__(self, n: int, seed: int = 0) -> None:
        self.num_envs = n
        self.seed = seed
        self.boards = np.zeros((n, 9), dtype=np.int8)
        self.agent_mark = np.ones(n, dtype=np.int8)

    def _reset_rows(self, rows: np.ndarray, u: np.ndarray) -> None:
        """``TicTacToeEnv.reset`` for the envs in ``rows`` (TTT:96-108)."""
        if rows.size == 0:
            return
        self.boards[rows] = 0
        starts = pick(u[rows, SLOT_ENV1], 2) == 0  # choice([True, False]): index 0 is True
        self.agent_mark[rows] = np.where(starts, 1, 2)
        opener = rows[~starts]
        self.boards[opener, pick(u[opener, SLOT_ENV2], 9)] = 1  # machine (mark 1) opens

    def _obs(self):
        return {
            "observation": ttt_encode(self.boards),
            "action_mask": (self.boards == 0).astype(np.int64),  # TTT:213
        }

    def reset(self, uniforms: np.ndarray | None = None, seed=None, options=None):
        if uniforms is None:
            uniforms = draw_uniforms(self.seed, T_INIT, 1, self.num_envs, self.slots)[0]
        self._reset_rows(np.arange(self.num_envs), uniforms)
        return self._obs(), {}

    def step(self, actions: np.ndarray, uniforms: np.ndarray):
        n = self.num_envs
        ar = np.arange(n)
        actions = np.asarray(actions).astype(np.int64)
        assert (actions >= 0).all() and (actions < 9).all(), "Invalid move."
        assert (self.boards[ar, actions] == 0).all(), "Invalid move."  # TTT:130
        amark = self.agent_mark
        mmark = (3 - amark).astype(np.int8)
        self.boards[ar, actions] = amark
        win_a = ttt_has_line(self.boards, amark)
        full = (self.boards != 0).all(axis=1)
        term = win_a | full  # TTT:132-136
        rewards = np.where(win_a, 1.0, 0.0).astype(np.float32)
        live = np.nonzero(~term)[0]
        if live.size:  # machine move: k-th empty cell ascending (TTT:183-197)
            empty = self.boards[live] == 0
            cnt = empty.sum(axis=1)
            k = pick(uniforms[live, SLOT_ENV0], cnt)
            rank = np.cumsum(empty, axis=1) - 1
            cell = (empty & (rank == k[:, None])).argmax(axis=1)
            self.boards[live, cell] = mmark[live]
            win_m = ttt_has_line(self.boards[live], mmark[live])
            full2 = (self.boards[live] != 0).all(axis=1)
            rewards[live[win_m]] = -1.0
            term[live] = win_m | full2
        self._reset_rows(np.nonzero(term)[0], uniforms)  # SAME_STEP autoreset
        return self._obs(), rewards, term.copy(), np.zeros(n, dtype=bool), {}


class BanditVec:
    num_actions = 2
    num_states = 1
    slots = 2

    def __init__(self, n: int, episode_len: int = 10) -> None:
        self.num_envs = n
        self.episode_len = episode_len
        self.t = np.zeros(n, dtype=np.int64)

    def reset(self, uniforms=None, seed=None, options=None):
        self.t[:] = 0
        return np.zeros(self.num_envs, dtype=np.int64), [{} for _ in range(self.num_envs)]

    def step(self, actions, uniforms=None):
        actions = np.asarray(actions).astype(np.int64)
        assert ((actions == 0) | (actions == 1)).all(), "Invalid action"
        self.t += 1
        term = self.t >= self.episode_len
        self.t[term] = 0
        n = self.num_envs
        return (
            np.zeros(n, dtype=np.int64),
            actions.astype(np.float32),
            term,
            np.zeros(n, dtype=bool),
            [{} for _ in range(n)],
        )


SALT_TRANSITION = 0
SALT_REWARD = 1
SALT_MASK = 2


def mdp_mask_bits(states: np.ndarray, num_actions: int, seed: int) -> np.ndarray:
    """Legal-action bitmask of a state: ``mix32(s ^ seedmix, 2) | 1`` cut to ``A`` bits."""
    h = mix32(np.asarray(states, dtype=np.uint64) + np.uint64((seed * 0x632BE5AB) & 0xFFFFFFFF), SALT_MASK)
    full = (1 << num_actions) - 1
    return (h.astype(np.uint64) & np.uint64(full)).astype(np.uint32) | np.uint32(1)


def bits_to_mask(bits: np.ndarray, num_actions: int) -> np.ndarray:
    return ((bits[:, None].astype(np.uint64) >> np.arange(num_actions, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(
        np.int64
    )


class HashMDPVec:
    """Synthetic tabular MDP defined by integer hashing (deterministic transitions).

    ``h = mix32(s*A + a + seedmix, 0)``; ``s' = (h*S) >> 32``;
    ``r = float32((mix32(h, 1) >> 8) * 2**-24) * 2 - 1``;
    ``terminated = U[t,i,2] < ceil(p_term * 2**32)``; on termination the agent
    restarts (SAME_STEP) in ``(U[t,i,3]*S) >> 32``.  Requires ``S*A < 2**32``.
    """

    slots = 4

    def __init__(self, n: int, num_states: int, num_actions: int, seed: int = 0, p_term: float = 0.05) -> None:
        assert num_states * num_actions < 2**32 and num_actions <= 32
        self.num_envs = n
        self.num_states = num_states
        self.num_actions = num_actions
        self.seed = seed
        self.p_term = p_term
        self.term_threshold = int(np.ceil(p_term * 2.0**32))
        self.states = np.zeros(n, dtype=np.int64)

    def _obs(self):
        return {
            "observation": self.states.copy(),
            "action_mask": bits_to_mask(mdp_mask_bits(self.states, self.num_actions, self.seed), self.num_actions),
        }

    def reset(self, uniforms: np.ndarray | None = None, seed=None, options=None):
        if uniforms is None:
            uniforms = draw_uniforms(self.seed, T_INIT, 1, self.num_envs, self.slots)[0]
        self.states = pick(uniforms[:, SLOT_ENV1], self.num_states)
        return self._obs(), {}

    def step(self, actions, uniforms: np.ndarray):
        actions = np.asarray(actions).astype(np.int64)
        assert (actions >= 0).all() and (actions < self.num_actions).all(), "Invalid action"
        seedmix = np.uint64((self.seed * 0x632BE5AB) & 0xFFFFFFFF)
        x = (self.states.astype(np.uint64) * np.uint64(self.num_actions) + actions.astype(np.uint64) + seedmix) & np.uint64(
            0xFFFFFFFF
        )
        h = mix32(x, SALT_TRANSITION)
        nxt = pick(h, self.num_states)
        h2 = mix32(h, SALT_REWARD)
        rewards = ((h2 >> np.uint32(8)).astype(np.float32) * np.float32(2.0**-24)) * np.float32(2.0) - np.float32(1.0)
        term = uniforms[:, SLOT_ENV0].astype(np.uint64) < np.uint64(self.term_threshold)
        restart = pick(uniforms[:, SLOT_ENV1], self.num_states)
        self.states = np.where(term, restart, nxt)
        n = self.num_envs
        return self._obs(), rewards.astype(np.float32), term, np.zeros(n, dtype=bool), {}
