"""ctypes binding of ``libqe_b200.so`` (the C ABI declared in ``include/qe_engine.h``).

There is no CPU fallback: if the shared library is missing and cannot be built, importing
the engine fails loudly.
"""

from __future__ import annotations

import ctypes as C
import os

from dist_classicrl_b200 import build as _build

QE_ABI_VERSION = 2  # include/qe_engine.h; checked against the library at load
QE_OK = 0
QE_ERR_ARG, QE_ERR_CUDA, QE_ERR_INVALID_MOVE, QE_ERR_EMPTY, QE_ERR_TIMEOUT = -1, -2, -3, -4, -5
QE_ENV_MDP, QE_ENV_TTT, QE_ENV_BANDIT = 0, 1, 2
QE_LEARN_SEQUENTIAL, QE_LEARN_ACCUMULATE = 0, 1

vp, i32, i64, u32, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float


class QeAgents(C.Structure):
    _fields_ = [
        ("env_kind", i32), ("num_agents", i32), ("states", vp), ("states_scratch", vp), ("env_words", vp),
        ("episode_returns", vp), ("env_seed", u32), ("episode_len", u32), ("term_threshold", u64),
    ]


class QeRun(C.Structure):
    _fields_ = [
        ("steps", i32), ("explore_thresholds_host", vp), ("learning_rates_host", vp), ("uniforms", vp), ("slots", i32),
        ("stream_seed", u32), ("t0", u32), ("agent0", u32), ("env_stream_seed", u32), ("env_t0", u32),
        ("empty_all", i32), ("use_masks", i32),
        ("trace_actions", vp), ("trace_rewards", vp), ("trace_terminated", vp), ("trace_next_states", vp),
        ("trace_episode_returns", vp), ("episode_sum", vp), ("episode_count", vp), ("evaluate", i32), ("learn_mode", i32),
    ]


# name -> (restype, argtypes); every symbol include/qe_engine.h declares
SIGNATURES = {
    "qe_create": (C.c_int, [i64, i32, f32, i32, C.POINTER(vp)]),
    "qe_destroy": (C.c_int, [vp]),
    "qe_abi_version": (C.c_int, []),
    "qe_last_error": (C.c_char_p, []),
    "qe_set_last_error": (C.c_int, [C.c_int, C.c_char_p]),
    "qe_shard_slab_bytes": (i64, [i64, i32, i32, i32]),
    "qe_shard_create": (C.c_int, [i64, i32, f32, i32, i32, i32, i32, u32, vp, C.POINTER(vp)]),
    "qe_shard_connect_ptr": (C.c_int, [vp, i32, vp]),
    "qe_shard_destroy": (C.c_int, [vp]),
    "qe_shard_ipc_handle": (C.c_int, [vp, vp]),
    "qe_shard_connect_ipc": (C.c_int, [vp, i32, vp]),
    "qe_shard_connect_local": (C.c_int, [vp, i32, vp]),
    "qe_shard_fill_random": (C.c_int, [vp, u32, vp]),
    "qe_shard_reset": (C.c_int, [vp, u32, u32, vp]),
    "qe_shard_steps": (C.c_int, [vp, i32, i32, vp, vp, u32, u32, i32, i32, u64, vp]),
    "qe_shard_sync": (C.c_int, [vp, vp]),
    "qe_shard_download": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "qe_shard_rows_host": (C.c_int, [vp, vp, vp, i32]),
    "qe_shard_phase_ns": (C.c_int, [vp, vp]),
    "qe_shard_probe": (C.c_double, [vp, i32, i32]),
    "qe_shard_info": (i32, [vp, i32]),
    "qe_set_discount": (C.c_int, [vp, f32]),
    "qe_table_ptr": (vp, [vp]),
    "qe_table_stride": (i32, [vp]),
    "qe_table_upload_host": (C.c_int, [vp, vp]),
    "qe_table_download_host": (C.c_int, [vp, vp]),
    "qe_table_fill": (C.c_int, [vp, f32, vp]),
    "qe_table_fill_random": (C.c_int, [vp, u32, vp]),
    "qe_sync": (C.c_int, [vp, vp]),
    "qe_radix_encode": (C.c_int, [vp, vp, i32, vp, i64, vp]),
    "qe_radix_decode": (C.c_int, [vp, vp, vp, i32, vp, i64, vp]),
    "qe_host_register": (C.c_int, [vp, u64]),
    "qe_host_unregister": (C.c_int, [vp]),
    "qe_select": (C.c_int, [vp, vp, vp, vp, vp, i32, u32, u32, u32, u64, i32, i32, vp, i32, vp]),
    "qe_select_host": (C.c_int, [vp, vp, vp, vp, vp, i32, u32, u32, u32, u64, i32, i32, vp, i32]),
    "qe_learn": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, f32, i32, i32, vp]),
    "qe_learn_host": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, f32, i32, i32]),
    "qe_gather": (C.c_int, [vp, vp, vp, vp, i32, vp]),
    "qe_gather_rows_host": (C.c_int, [vp, vp, vp, i32]),
    "qe_ttt_reset": (C.c_int, [vp, vp, vp, vp, i32, u32, u32, u32, i32, vp]),
    "qe_ttt_step": (C.c_int, [vp, vp, vp, vp, i32, u32, u32, u32, vp, vp, vp, vp, i32, vp]),
    "qe_mdp_reset": (C.c_int, [vp, vp, i64, i32, u32, vp, i32, u32, u32, u32, i32, vp]),
    "qe_mdp_masks": (C.c_int, [vp, vp, i32, u32, i32, vp]),
    "qe_mdp_step": (C.c_int, [vp, vp, vp, i64, i32, u32, u64, vp, i32, u32, u32, u32, vp, vp, vp, i32, vp]),
    "qe_fused_steps": (C.c_int, [vp, C.POINTER(QeAgents), C.POINTER(QeRun), vp]),
    "qe_set_state_base": (C.c_int, [vp, i64]),
    "qe_set_agent_ids": (C.c_int, [vp, vp]),
    "qe_set_hold": (C.c_int, [vp, i32]),
    "qe_learn_commit": (C.c_int, [vp, vp, vp, i32, vp]),
    "qe_serve_bootstrap": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, vp]),
    "qe_table_export_dense": (C.c_int, [vp, vp, vp]),
    "qe_table_import_dense": (C.c_int, [vp, vp, vp]),
    "qe_table_delta_dense": (C.c_int, [vp, vp, vp, vp]),
    "qe_table_merge_dense": (C.c_int, [vp, vp, vp, vp]),
    "qe_stream_u32": (u32, [u32, u32, u32, u32]),
    "qe_kernel_launches": (i64, [vp]),
    "qe_fused_grid_blocks": (i32, [vp]),
    "qe_fused_form": (i32, [vp]),
    "qe_set_fused_form": (C.c_int, [vp, i32]),
    "qe_fused_phase_ns": (i32, [vp, vp, i32]),
    "qe_debug_gather_gbs": (C.c_double, [vp]),
    "qe_debug_gridsync_us": (C.c_double, [vp, i32]),
    "qe_debug_counters": (C.c_int, [vp, vp, i32]),
    "qe_build_info": (C.c_char_p, []),
}

_lib = None


class EngineError(RuntimeError):
    pass


def library_path() -> str:
    return _build.OUT


def lib():
    """Load (building in-tree first if needed) the CUDA library.  Raises if it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("QE_LIBRARY") or _build.OUT  # QE_LIBRARY: development override (kernel variants)
    if path == _build.OUT and not _build.up_to_date():
        try:
            _build.build()
        except Exception as exc:  # noqa: BLE001
            # no stale binary behind newer sources: its struct layouts and signatures may not be the ones bound below
            raise ImportError(
                f"libqe_b200.so is missing or older than its sources and could not be rebuilt ({exc}); the B200 engine has no CPU fallback"
            ) from exc
    handle = C.CDLL(path)
    try:
        got = int(handle.qe_abi_version())
    except AttributeError:
        got = -1
    if got != QE_ABI_VERSION:
        raise ImportError(f"{path} was built for ABI version {got}, this binding expects {QE_ABI_VERSION}: rebuild with python -m dist_classicrl_b200.build --force")
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(rc: int) -> None:
    """Map C-ABI status codes onto the exceptions the reference raises for the same condition."""
    if rc == QE_OK:
        return
    msg = lib().qe_last_error().decode(errors="replace")
    if rc == QE_ERR_INVALID_MOVE:
        raise AssertionError(msg or "Invalid move.")  # TTT:130
    if rc == QE_ERR_EMPTY:
        raise IndexError(msg)  # choice([]) / np.max of empty (QLO:470, 764)
    if rc == QE_ERR_ARG:
        raise ValueError(msg)
    raise EngineError(f"qe error {rc}: {msg}")
