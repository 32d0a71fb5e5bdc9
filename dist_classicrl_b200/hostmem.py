"""Page-locking of caller-owned host arrays (so that the runtimes' per-step copies are asynchronous DMA).

The reference's state dictionaries hold plain NumPy arrays (``STR:57, 70-75``) that the trainer updates in place.
The fused path has to move such an array (the agents' running returns, 4 B per agent) to the device before a launch
and back after it; from pageable memory both copies are staged by the driver at a few GB/s.  ``pin`` page-locks
the array's own memory with ``cudaHostRegister`` -- the array object, its address and its contents stay what the
caller handed in -- and un-registers it when the array object dies.
"""

from __future__ import annotations

import weakref

import numpy as np

_registered: dict[int, int] = {}  # address -> bytes (ranges registered by this module)


def _unpin(ptr: int) -> None:
    if _registered.pop(ptr, None) is not None:
        try:
            from dist_classicrl_b200 import capi

            capi.lib().qe_host_unregister(ptr)
        except Exception:  # noqa: BLE001  (interpreter shutdown: the driver releases the mapping with the context)
            pass


def pin(arr) -> bool:
    """Page-lock ``arr`` in place.  True if its memory is page-locked afterwards (registered now or earlier, or
    allocated pinned by somebody else), False if it cannot be registered (copies then stay synchronous)."""
    from dist_classicrl_b200 import capi

    if not isinstance(arr, np.ndarray) or not arr.flags.c_contiguous or not arr.flags.writeable or arr.nbytes == 0:
        return False
    ptr = arr.ctypes.data
    if _registered.get(ptr, 0) >= arr.nbytes:
        return True
    if ptr in _registered:  # registered with a shorter length: start over
        _unpin(ptr)
    rc = capi.lib().qe_host_register(ptr, arr.nbytes)
    if rc < 0:
        return False
    if rc == 1:
        _registered[ptr] = arr.nbytes
        weakref.finalize(arr, _unpin, ptr)
    return True
