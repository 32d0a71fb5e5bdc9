"""CPU restatement of the reference's mixed-radix utilities (TEST INFRASTRUCTURE -- only tests/ may import this).

Follows ``/root/reference/src/dist_classicrl/utils.py``: ``compute_radix`` :12-29, ``encode_multi_discrete`` :32-48,
``encode_multi_discretes`` :51-69, ``decode_to_multi_discrete`` :72-92, ``decode_to_multi_discretes`` :95-115.
Pinned by ``tests/golden/radix.npz`` (outputs of the live reference, ``oracle/make_golden_radix.py``).
"""

from __future__ import annotations

import numpy as np


def compute_radix(nvec):  # UTL:26-29
    nvec = np.asarray(nvec)
    shifted = np.concatenate([[1], nvec[::-1][:-1]])
    return np.cumprod(shifted, dtype=np.int32)[::-1]


def encode(vectors, radix):  # UTL:48 (one vector: dot), UTL:69 (batch: sum of products along axis 1)
    vectors = np.asarray(vectors)
    if vectors.ndim == 1:
        return int(np.dot(vectors, radix))
    return np.sum(vectors * radix, axis=1)


def decode(nvec, indices, radix):  # UTL:92, 115: floor division then modulo, broadcast over the dims
    indices = np.asarray(indices)
    if indices.ndim == 0:
        return (indices // radix) % nvec
    return (indices.reshape(-1, 1) // np.asarray(radix)[None, :]) % np.asarray(nvec)[None, :]
