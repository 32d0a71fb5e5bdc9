from __future__ import annotations

import numpy as np


class Space:
    pass


class Discrete(Space):
    def __init__(self, n, start=0):
        self.n = int(n)
        self.start = int(start)
        self.dtype = np.int64

    def contains(self, x):
        try:
            xi = int(x)
        except (TypeError, ValueError):
            return False
        return xi == x and self.start <= xi < self.start + self.n


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.dtype = np.int64


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)
