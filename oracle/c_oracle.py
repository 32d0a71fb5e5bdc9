"""ctypes binding of ``oracle/c/liboracle.so`` (TEST INFRASTRUCTURE; see ``oracle/c/oracle.c``)."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_DIR, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-s", "-B", "liboracle.so"])
    return _SO


class _Trace(C.Structure):
    _fields_ = [("actions", C.c_void_p), ("rewards", C.c_void_p), ("term", C.c_void_p),
                ("next_states", C.c_void_p), ("episode_returns", C.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_stream_u32.restype = C.c_uint32
        _lib.orc_stream_u32.argtypes = [C.c_uint32] * 4
        _lib.orc_run.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def masks_to_bits(masks) -> np.ndarray | None:
    """``[N, A]`` truthy array -> ``uint32 [N]`` bitmask (bit a = action a legal), A <= 32."""
    if masks is None:
        return None
    m = np.asarray(masks).astype(bool)
    assert m.shape[1] <= 32
    w = (np.uint64(1) << np.arange(m.shape[1], dtype=np.uint64))[None, :]
    return (m * w).sum(axis=1).astype(np.uint32)


def select(q, states, mask_bits, explore_thresh, deterministic, empty_all, u):
    q = np.ascontiguousarray(q, dtype=np.float32)
    states = np.ascontiguousarray(states, dtype=np.int32)
    u = np.ascontiguousarray(u, dtype=np.uint32)
    n = states.shape[0]
    out = np.empty(n, dtype=np.int32)
    lib().orc_select(_p(q), C.c_int(q.shape[1]), _p(states), _p(mask_bits), C.c_uint64(explore_thresh),
                     C.c_int(int(deterministic)), C.c_int(int(empty_all)), _p(u), C.c_int(u.shape[1]), C.c_int(n), _p(out))
    return out


def learn_seq(q, s, a, r, s2, term, mask_bits2, lr, gamma):
    assert q.dtype == np.float32 and q.flags.c_contiguous
    s = np.ascontiguousarray(s, dtype=np.int32)
    a = np.ascontiguousarray(a, dtype=np.int32)
    r = np.ascontiguousarray(r, dtype=np.float32)
    s2 = np.ascontiguousarray(s2, dtype=np.int32)
    term = np.ascontiguousarray(term, dtype=np.uint8)
    lib().orc_learn_seq(_p(q), C.c_int(q.shape[1]), _p(s), _p(a), _p(r), _p(s2), _p(term), _p(mask_bits2),
                        C.c_float(np.float32(lr)), C.c_float(np.float32(gamma)), C.c_int(s.shape[0]))


def ttt_reset(u_init):
    u = np.ascontiguousarray(u_init, dtype=np.uint32)
    n = u.shape[0]
    boards, states, masks = np.empty(n, np.uint32), np.empty(n, np.int32), np.empty(n, np.uint32)
    lib().orc_ttt_reset(_p(boards), _p(states), _p(masks), _p(u), C.c_int(u.shape[1]), C.c_int(n))
    return boards, states, masks


def mdp_reset(u_init, num_states, num_actions, env_seed):
    u = np.ascontiguousarray(u_init, dtype=np.uint32)
    n = u.shape[0]
    states, masks = np.empty(n, np.int32), np.empty(n, np.uint32)
    lib().orc_mdp_reset(_p(states), C.c_int(num_states), _p(u), C.c_int(u.shape[1]), C.c_int(n))
    lib().orc_mdp_masks(_p(states), C.c_int(num_actions), C.c_uint32(env_seed), C.c_int(n), _p(masks))
    return states, masks


ENV_MDP, ENV_TTT = 0, 1


def run(env_kind, q, env_state, states, masks, *, num_states, env_seed=0, term_thresh=0, uniforms=None, slots=4,
        stream_seed=0, t0=0, agent0=0, steps, eps_thresh, lr, gamma, empty_all=False, agent_rewards=None, record=False):
    """Fused CPU loop (select -> step -> learn) for ``steps`` vector steps; arrays are updated in place.

    Returns ``dict(rc, ep_sum, ep_count, trace)``.
    """
    assert q.dtype == np.float32 and q.flags.c_contiguous
    n = states.shape[0]
    if agent_rewards is None:
        agent_rewards = np.zeros(n, dtype=np.float32)
    eps_thresh = np.ascontiguousarray(eps_thresh, dtype=np.uint64)
    lr = np.ascontiguousarray(lr, dtype=np.float32)
    assert eps_thresh.shape[0] >= steps and lr.shape[0] >= steps
    ep_sum, ep_count = C.c_double(0.0), C.c_int64(0)
    trace, tr = None, None
    if record:
        trace = {"actions": np.empty((steps, n), np.int32), "rewards": np.empty((steps, n), np.float32),
                 "terminated": np.empty((steps, n), np.uint8), "obs": np.empty((steps, n), np.int32),
                 "episode_returns": np.empty((steps, n), np.float32)}
        tr = _Trace(_p(trace["actions"]), _p(trace["rewards"]), _p(trace["terminated"]), _p(trace["obs"]),
                    _p(trace["episode_returns"]))
    if uniforms is not None:
        uniforms = np.ascontiguousarray(uniforms, dtype=np.uint32)
        slots = uniforms.shape[2]
    rc = lib().orc_run(C.c_int(env_kind), _p(q), C.c_int(num_states), C.c_int(q.shape[1]), C.c_int(n), _p(env_state),
                       _p(states), _p(masks), C.c_uint32(env_seed), C.c_uint64(term_thresh), _p(uniforms), C.c_int(slots),
                       C.c_uint32(stream_seed), C.c_uint32(t0), C.c_uint32(agent0), C.c_int(steps), _p(eps_thresh), _p(lr),
                       C.c_float(np.float32(gamma)), C.c_int(int(empty_all)), _p(agent_rewards), C.byref(ep_sum),
                       C.byref(ep_count), C.byref(tr) if tr is not None else None)
    return {"rc": rc, "ep_sum": ep_sum.value, "ep_count": ep_count.value, "trace": trace, "agent_rewards": agent_rewards}
