"""NumPy restatement of ``OptimalQLearningBase`` select / learn (TEST INFRASTRUCTURE).

Reference: ``src/dist_classicrl/algorithms/base_algorithms/q_learning_optimal.py`` (QLO).
All eight ``choose_*`` variants (QLO:263-642) implement one function of
``(Q, states, masks, eps, U[t])`` -- they differ only in which RNG method they
call -- so the oracle has ONE select; the single behavioural difference (the
all-zero mask edge case, QLO:348 vs QLO:467-470) is the ``empty`` argument.

fp32 pipeline: with ``q_table.dtype == float32`` and float32
rewards NumPy-2 promotion makes ``single_learn`` (QLO:758-768) compute
``q + f32(lr) * ((r + f32(gamma) * m) - q)`` with a rounding after every
operation; :func:`learn_sequential` reproduces exactly that via NumPy scalars.
"""

from __future__ import annotations

import numpy as np

from oracle.rng import SLOT_EXPLORE, SLOT_PICK, explore_threshold

# dispatcher thresholds, QLO:14-20
DETERMINISTIC_MAX_ACTION_SIZE_ITER = 10
DETERMINISTIC_MIN_ACTION_SIZE_VEC_ITER = 10000
DETERMINISTIC_MAX_NUM_STATES_VEC_ITER = 3
ACTION_MASKS_NO_DETERMINISTIC_MAX_ACTION_SIZE_ITER = 10

EMPTY_MINUS1 = "minus1"  # choose_masked_action (QLO:348): no candidate -> -1
EMPTY_ALL = "all"  # choose_masked_action_vec (QLO:467-470): -inf ties -> every action


def empty_policy(action_size: int, n_states: int, deterministic: bool) -> str:
    """Which all-zero-mask behaviour the dispatcher (QLO:644-726) reaches."""
    if deterministic:
        if action_size <= DETERMINISTIC_MAX_ACTION_SIZE_ITER:
            return EMPTY_MINUS1
        return EMPTY_ALL
    if action_size <= ACTION_MASKS_NO_DETERMINISTIC_MAX_ACTION_SIZE_ITER:
        return EMPTY_MINUS1
    return EMPTY_ALL


def select(
    q: np.ndarray,
    states: np.ndarray,
    masks: np.ndarray | None,
    eps: float,
    uniforms_t: np.ndarray,
    *,
    deterministic: bool = False,
    empty: str | None = None,
) -> np.ndarray:
    """Masked epsilon-greedy selection for one vector step (phase 1).

    ``uniforms_t`` is ``U[t]`` (``uint32 [N, >=2]``).  Returns ``int32 [N]``.
    """
    states = np.asarray(states).astype(np.int64)
    n = states.shape[0]
    a_size = q.shape[1]
    if empty is None:
        empty = empty_policy(a_size, n, deterministic)
    valid = np.ones((n, a_size), dtype=bool) if masks is None else np.asarray(masks).astype(bool)
    assert valid.shape == (n, a_size), "Action masks must match the number of states and actions."
    bits0 = uniforms_t[:, SLOT_EXPLORE].astype(np.uint64)
    bits1 = uniforms_t[:, SLOT_PICK].astype(np.uint64)
    if deterministic:
        explore = np.zeros(n, dtype=bool)  # QLO:287,335: `not deterministic and ...`
    else:
        explore = bits0 < np.uint64(explore_threshold(eps))  # u < eps, strict (QLO:335,464)
    rows = q[states]  # [N, A]
    masked = np.where(valid, rows, -np.inf)  # QLO:467,613 (-inf fill)
    best = masked.max(axis=1, keepdims=True)
    ties = masked == best  # exact == on the table dtype (QLO:346,469)
    if empty == EMPTY_MINUS1:
        ties &= valid
    cand = np.where(explore[:, None], valid, ties)  # ascending action order
    cnt = cand.sum(axis=1).astype(np.uint64)
    idx = ((bits1 * cnt) >> np.uint64(32)).astype(np.int64)  # choice(cand) / randint
    rank = np.cumsum(cand, axis=1) - 1
    hit = cand & (rank == idx[:, None])
    actions = np.where(cnt > 0, hit.argmax(axis=1), -1)
    return actions.astype(np.int32)


def learn_sequential(
    q: np.ndarray,
    states,
    actions,
    rewards,
    next_states,
    terminated,
    lr: float,
    gamma: float,
    next_masks=None,
) -> None:
    """``learn`` -> ``learn_iter`` -> ``single_learn`` (QLO:893-934, 770-817, 728-768).

    Strictly sequential in agent order, in place.  With a float32 table and float32
    rewards every operation is a float32 operation (rounded after each one).
    """
    rewards = np.asarray(rewards)  # dtype decides fp32 vs fp64 evaluation (NEP 50 promotion)
    lr, gamma = float(lr), float(gamma)  # weak python scalars, as in the reference
    for i in range(len(states)):
        s, a, s2 = int(states[i]), int(actions[i]), int(next_states[i])
        if terminated[i]:
            m = 0  # QLO:759,762-763
        elif next_masks is None:
            m = q[s2].max()  # QLO:759
        else:
            m = q[s2][np.nonzero(next_masks[i])[0]].max()  # QLO:764 (raises on empty, like np.max)
        target = rewards[i] + gamma * m  # QLO:766
        p = q[s, a]  # QLO:767
        q[s, a] += lr * (target - p)  # QLO:768, 233


def learn_accumulate(
    q: np.ndarray,
    states,
    actions,
    rewards,
    next_states,
    terminated,
    lr: float,
    gamma: float,
    next_masks=None,
) -> None:
    """``_learn_vec`` (QLO:853-891): snapshot bootstrap + accumulating scatter (``np.add.at``)."""
    states = np.asarray(states).astype(np.int64)
    actions = np.asarray(actions).astype(np.int64)
    next_states = np.asarray(next_states).astype(np.int64)
    rows = q[next_states]
    if next_masks is None:
        mx = rows.max(axis=1)
    else:
        mx = np.where(np.asarray(next_masks).astype(bool), rows, -np.inf).max(axis=1)
    targets = np.asarray(rewards) + gamma * mx * (1 - np.asarray(terminated).astype(np.int64))
    pred = q[states, actions]
    np.add.at(q, (states, actions), lr * (targets - pred))
