"""B200-native engine for the select -> env step -> TD update hot path of ``dist_classicrl``.

The package mirrors the reference's module layout for that path (``algorithms/base_algorithms``,
``algorithms/runtime``, ``environments``, ``schedules``); the work itself runs in hand-written sm_100a CUDA
kernels behind the C ABI of ``include/qe_engine.h`` (``dist_classicrl_b200/csrc``).  There is no CPU fallback.
"""

__version__ = "0.1.0"
