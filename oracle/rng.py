"""Shared integer hash, pre-drawn uniform stream and RNG shims (TEST INFRASTRUCTURE).

"Identical inputs and pre-drawn uniform random numbers" (BASELINE.json) is made
concrete here.  Every random decision on the hot path consumes one ``uint32``
``U[t, i, k]`` (vector step ``t``, agent ``i``, slot ``k``):

====  =========================================================  =====================
slot  reference call site                                        meaning
====  =========================================================  =====================
0     ``_rng.uniform(0,1)`` QLO:287,335 / ``_rng.random()``      explore test
      QLO:426,464 / ``_np_rng.random(N)`` QLO:551,617
1     ``_rng.choice(cand)`` QLO:300,348,430,470,556,621 /        pick among candidates
      ``_rng.randint(0,A-1)`` QLO:288,427 /
      ``_np_rng.integers(A,size=N)`` QLO:552
2     TicTacToe ``_np_random.choice(valid_moves)`` TTT:185       machine move
      | MDP: termination draw
3     TicTacToe ``_np_random.choice([True, False])`` TTT:98      who starts after reset
      | MDP: reset state
4     TicTacToe ``_np_random.choice(range(9))`` TTT:106          machine opening move
====  =========================================================  =====================

``u = bits * 2**-32`` (exact in fp64) and ``pick(bits, n) = (bits * n) >> 32``.
The CUDA kernels (``dist_classicrl_b200/csrc/qe_common.cuh``) and the C oracle
(``oracle/c/oracle.c``) implement the same formulas; the stream itself is the
counter hash :func:`stream_u32`, so "pre-drawn" arrays and on-device
generation give identical numbers.
"""

from __future__ import annotations

import numpy as np

M32 = 0xFFFFFFFF
GOLD = 0x9E3779B9
STREAM_ADD = 0x7F4A7C15
MAX_SLOTS = 8  # U[t, i, 0:8]

SLOT_EXPLORE = 0
SLOT_PICK = 1
SLOT_ENV0 = 2
SLOT_ENV1 = 3
SLOT_ENV2 = 4


def fmix32(x):
    """murmur3 32-bit finaliser; works on python ints and on uint32/uint64 arrays."""
    if isinstance(x, np.ndarray):
        x = x.astype(np.uint64) & M32
        x ^= x >> np.uint64(16)
        x = (x * np.uint64(0x85EBCA6B)) & np.uint64(M32)
        x ^= x >> np.uint64(13)
        x = (x * np.uint64(0xC2B2AE35)) & np.uint64(M32)
        x ^= x >> np.uint64(16)
        return x.astype(np.uint32)
    x &= M32
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & M32
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & M32
    x ^= x >> 16
    return x


def mix32(x, salt: int):
    """Salted avalanche hash ``fmix32(x + GOLD*(salt+1))`` (mod 2**32)."""
    add = (GOLD * (int(salt) + 1)) & M32
    if isinstance(x, np.ndarray):
        return fmix32((x.astype(np.uint64) + np.uint64(add)) & np.uint64(M32))
    return fmix32((int(x) + add) & M32)


def stream_u32(seed: int, t, i, k):
    """``U[t, i, k]`` of the counter-based uniform stream (``i < 2**29``, ``k < 8``)."""
    s1 = (int(seed) * GOLD) & M32
    if isinstance(i, np.ndarray) or isinstance(t, np.ndarray) or isinstance(k, np.ndarray):
        i = np.asarray(i, dtype=np.uint64)
        k = np.asarray(k, dtype=np.uint64)
        t = np.asarray(t, dtype=np.uint64)
        a = ((i * np.uint64(8) + k) & np.uint64(M32)) ^ np.uint64(s1)
        inner = fmix32(a).astype(np.uint64)
        return fmix32((inner + t * np.uint64(GOLD) + np.uint64(STREAM_ADD)) & np.uint64(M32))
    a = ((int(i) * 8 + int(k)) & M32) ^ s1
    return fmix32((fmix32(a) + int(t) * GOLD + STREAM_ADD) & M32)


def draw_uniforms(seed: int, t0: int, steps: int, n_agents: int, slots: int, agent0: int = 0):
    """Materialise ``U[t0:t0+steps, agent0:agent0+n_agents, 0:slots]`` as ``uint32``."""
    t = np.arange(t0, t0 + steps, dtype=np.uint64)[:, None, None]
    i = np.arange(agent0, agent0 + n_agents, dtype=np.uint64)[None, :, None]
    k = np.arange(slots, dtype=np.uint64)[None, None, :]
    return stream_u32(seed, t, i, k).reshape(steps, n_agents, slots)


def pick(bits, n):
    """``(bits * n) >> 32`` -- index in ``[0, n)`` from 32 random bits."""
    if isinstance(bits, np.ndarray) or isinstance(n, np.ndarray):
        return ((np.asarray(bits, dtype=np.uint64) * np.asarray(n, dtype=np.uint64)) >> np.uint64(32)).astype(
            np.int64
        )
    return (int(bits) * int(n)) >> 32


def u01(bits):
    """``bits * 2**-32`` as float64 (exact)."""
    if isinstance(bits, np.ndarray):
        return bits.astype(np.float64) * 2.0**-32
    return int(bits) * 2.0**-32


def explore_threshold(eps: float) -> int:
    """Integer ``T`` with ``u01(bits) < eps  <=>  bits < T`` for every uint32 ``bits``.

    ``bits * 2**-32 < eps  <=>  bits < eps * 2**32`` (power-of-two scaling is exact),
    and an integer is below a real iff it is below its ceiling.  NaN never explores.
    """
    import math

    if eps != eps or eps <= 0.0:
        return 0
    x = eps * 4294967296.0
    if x >= 4294967296.0:
        return 1 << 32
    return int(math.ceil(x))


class PredrawnRNG:
    """Duck-typed replacement for ``OptimalQLearningBase._rng`` *and* ``._np_rng``.

    Precedent: the reference's own tests swap ``algo._rng`` for a shim
    (T-RT:17-45,70; T-MPI:36-91).  Call :meth:`begin_step` before every
    ``choose_actions`` call; the shim then serves slot 0 / slot 1 of the current
    agent in reference call order.
    """

    def __init__(self, uniforms: np.ndarray) -> None:
        assert uniforms.dtype == np.uint32 and uniforms.ndim == 3
        self.U = uniforms
        self.t = 0
        self.i = 0
        self.eps = 0.0
        self._awaiting_pick = False
        self._queue: list[int] | None = None

    def begin_step(self, t: int, eps: float = 0.0) -> None:
        self.t, self.i, self.eps = t, 0, eps
        self._awaiting_pick = False
        self._queue = None

    # --- random.Random surface -------------------------------------------------
    def _explore(self) -> float:
        if self._awaiting_pick:  # previous agent had no candidates (QLO:348 returns -1)
            self.i += 1
        self._awaiting_pick = True
        return u01(int(self.U[self.t, self.i, SLOT_EXPLORE]))

    def uniform(self, a: float = 0.0, b: float = 1.0) -> float:
        return a + (b - a) * self._explore()

    def random(self) -> float:
        return self._explore()

    def _pick(self, n: int) -> int:
        if self._queue is not None:
            i = self._queue.pop(0)
        else:
            i = self.i
            self.i += 1
            self._awaiting_pick = False
        return pick(int(self.U[self.t, i, SLOT_PICK]), n)

    def randint(self, a: int, b: int) -> int:
        return a + self._pick(b - a + 1)

    def choice(self, seq):
        return seq[self._pick(len(seq))]

    # --- numpy.random.Generator surface (batch variants QLO:551-552, 617) ------
    def _np_random(self, n: int) -> np.ndarray:
        u = u01(self.U[self.t, :n, SLOT_EXPLORE])
        # masked batch variant: ``choice`` is called for every agent in order;
        # unmasked variant (``integers`` called next): only for exploiters.
        self._queue = list(range(n))
        self._u = u
        return u

    def _np_integers(self, high: int, size: int) -> np.ndarray:
        self._queue = [i for i in range(size) if not (self._u[i] < self.eps)]
        return pick(self.U[self.t, :size, SLOT_PICK], high)


class _NpView:
    """``_np_rng`` view of a :class:`PredrawnRNG` (``random(n)``, ``integers(high, size=n)``)."""

    def __init__(self, parent: PredrawnRNG) -> None:
        self._p = parent

    def random(self, n: int) -> np.ndarray:
        return self._p._np_random(int(n))

    def integers(self, high: int, size: int) -> np.ndarray:
        return self._p._np_integers(int(high), int(size))


def install_predrawn(algo, uniforms: np.ndarray) -> PredrawnRNG:
    """Replace ``algo._rng`` / ``algo._np_rng`` (reference or engine class) by the shim."""
    shim = PredrawnRNG(uniforms)
    algo._rng = shim
    algo._np_rng = _NpView(shim)
    return shim
