import numpy as np


def np_random(seed=None):
    """``(Generator, seed)`` like gymnasium.utils.seeding.np_random."""
    ss = np.random.SeedSequence(seed)
    return np.random.Generator(np.random.PCG64(ss)), (ss.entropy if seed is None else seed)
