"""Tiny stand-ins for the ``gymnasium.spaces`` objects the reference inspects (TPB:133-143): gymnasium is a
third-party dependency of the reference and not a requirement of this engine."""

from __future__ import annotations

import numpy as np


class Discrete:
    def __init__(self, n: int, start: int = 0) -> None:
        self.n = int(n)
        self.start = int(start)
        self.dtype = np.int64

    def contains(self, x) -> bool:
        try:
            xi = int(x)
        except (TypeError, ValueError):
            return False
        return xi == x and self.start <= xi < self.start + self.n

    def __repr__(self) -> str:
        return f"Discrete({self.n})"


class MultiDiscrete:
    def __init__(self, nvec) -> None:
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.dtype = np.int64

    def __repr__(self) -> str:
        return f"MultiDiscrete({self.nvec.tolist()})"


class Dict:
    def __init__(self, spaces) -> None:
        self.spaces = dict(spaces)

    def __getitem__(self, key):
        return self.spaces[key]

    def __repr__(self) -> str:
        return f"Dict({self.spaces})"
