"""Uniform stream of the engine (host side) and the RNG objects that sit in the reference's ``_rng`` seam.

Every random decision on the hot path consumes one ``uint32`` ``U[t, i, k]`` (vector step ``t``, agent ``i``,
slot ``k``): slot 0 = explore test (``_rng.uniform(0,1)`` QLO:335 / ``_rng.random()`` QLO:464), slot 1 = pick
among the candidates (``_rng.choice`` QLO:348,470 / ``_rng.randint`` QLO:288), slots 2.. = environment draws
(TTT:98,106,185).  ``u = bits * 2**-32``; ``pick(bits, n) = (bits * n) >> 32``.  ``U`` is a counter hash
(``stream_u32``), so the CUDA kernels can generate it on the fly and a host array of "pre-drawn" numbers is
just the same function tabulated.
"""

from __future__ import annotations

import math
import os
import random

import numpy as np

M32 = 0xFFFFFFFF
GOLD = 0x9E3779B9
STREAM_ADD = 0x7F4A7C15
T_INIT = 0xFFFFFFFF  # stream index of the first environment reset (later resets count down from it)


def _fmix32(x: np.ndarray) -> np.ndarray:
    x = x & np.uint64(M32)
    x = x ^ (x >> np.uint64(16))
    x = (x * np.uint64(0x85EBCA6B)) & np.uint64(M32)
    x = x ^ (x >> np.uint64(13))
    x = (x * np.uint64(0xC2B2AE35)) & np.uint64(M32)
    return x ^ (x >> np.uint64(16))


def draw_uniforms(seed: int, t0: int, steps: int, n_agents: int, slots: int, agent0: int = 0) -> np.ndarray:
    """Tabulate ``U[t0:t0+steps, agent0:agent0+n_agents, 0:slots]`` (``uint32``)."""
    t = np.arange(t0, t0 + steps, dtype=np.uint64)[:, None, None]
    i = np.arange(agent0, agent0 + n_agents, dtype=np.uint64)[None, :, None]
    k = np.arange(slots, dtype=np.uint64)[None, None, :]
    a = ((i * np.uint64(8) + k) & np.uint64(M32)) ^ np.uint64((int(seed) * GOLD) & M32)
    inner = _fmix32(a)
    out = _fmix32((inner + t * np.uint64(GOLD) + np.uint64(STREAM_ADD)) & np.uint64(M32))
    return out.astype(np.uint32)


def explore_threshold(eps: float) -> int:
    """Integer ``T`` such that ``bits * 2**-32 < eps  <=>  bits < T`` for every ``uint32`` (strict ``<``, QLO:335)."""
    eps = float(eps)
    if eps != eps or eps <= 0.0:
        return 0
    x = eps * 4294967296.0
    return (1 << 32) if x >= 4294967296.0 else int(math.ceil(x))


class CounterRNG:
    """Default occupant of ``algo._rng`` / ``env._rng``: a seed plus a vector-step counter.

    The hot path never calls its scalar methods -- kernels evaluate ``U[t, i, k]`` themselves.  The
    ``random.Random``-style methods exist so that code which pokes the seam directly keeps working.
    """

    def __init__(self, seed: int | None = None) -> None:
        if seed is None:
            seed = int.from_bytes(os.urandom(4), "little")
        self.seed = int(seed) & M32
        self.t = 0
        self._scalar = random.Random(seed)

    def next_step(self) -> int:
        t = self.t
        self.t = (self.t + 1) & M32
        return t

    def uniform(self, a: float = 0.0, b: float = 1.0) -> float:
        return self._scalar.uniform(a, b)

    def random(self) -> float:
        return self._scalar.random()

    def randint(self, a: int, b: int) -> int:
        return self._scalar.randint(a, b)

    def choice(self, seq):
        return seq[self._scalar.randrange(len(seq))]


class PredrawnUniforms:
    """``_rng`` occupant that serves a caller-supplied array ``U[T, N, K]`` (``uint32``), one row per vector step."""

    def __init__(self, uniforms: np.ndarray, t0: int = 0) -> None:
        u = np.ascontiguousarray(uniforms, dtype=np.uint32)
        if u.ndim != 3:
            raise ValueError("uniforms must have shape [steps, agents, slots]")
        self.uniforms = u
        self.t = t0

    def next_step(self) -> int:
        t = self.t
        self.t += 1
        if t >= self.uniforms.shape[0]:
            raise IndexError("pre-drawn uniform stream exhausted")
        return t

    def row(self, t: int) -> np.ndarray:
        return self.uniforms[t]


_SCALAR_METHODS = ("uniform", "random", "randint", "choice")


def is_engine_rng(rng) -> bool:
    """True if ``rng`` is one of ours AND nobody patched its scalar methods (``unittest.mock.patch.object``)."""
    if type(rng) not in (CounterRNG, PredrawnUniforms):
        return False
    return not any(m in vars(rng) for m in _SCALAR_METHODS)
