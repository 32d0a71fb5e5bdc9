"""The C restatement (oracle/c/oracle.c) against the golden fixtures and the NumPy oracle."""
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import rng as orng
from oracle.envs import T_INIT, mdp_mask_bits
from oracle.runtime import Exponential

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
SELECT = np.load(os.path.join(GOLDEN, "select.npz"))
LEARN = np.load(os.path.join(GOLDEN, "learn.npz"))
TTT = np.load(os.path.join(GOLDEN, "ttt_traj.npz"))
MDP = np.load(os.path.join(GOLDEN, "mdp_traj.npz"))


def _cases(npz):
    return sorted({k.split("__")[0] for k in npz.files})


def test_stream_matches_numpy():
    u = orng.draw_uniforms(5, 3, 2, 7, 5)
    for t in range(2):
        for i in range(7):
            for k in range(5):
                assert co.lib().orc_stream_u32(5, 3 + t, i, k) == int(u[t, i, k])
    assert co.lib().orc_stream_u32(9, T_INIT, 123456, 3) == orng.stream_u32(9, T_INIT, 123456, 3)


@pytest.mark.parametrize("name", [c for c in _cases(SELECT) if SELECT[f"{c}__q"].shape[1] <= 32])
def test_select(name):
    g = lambda k: SELECT[f"{name}__{k}"]  # noqa: E731
    masks = g("masks") if f"{name}__masks" in SELECT.files else None
    a = g("q").shape[1]
    got = co.select(g("q"), g("states"), co.masks_to_bits(masks), orng.explore_threshold(float(g("eps"))),
                    bool(g("det")), a > 10, g("u"))
    np.testing.assert_array_equal(got, g("actions"))


@pytest.mark.parametrize("name", [c for c in _cases(LEARN) if LEARN[f"{c}__q0"].dtype == np.float32])
def test_learn(name):
    g = lambda k: LEARN[f"{name}__{k}"]  # noqa: E731
    masks = g("masks") if f"{name}__masks" in LEARN.files else None
    q = g("q0").copy()
    co.learn_seq(q, g("states"), g("actions"), g("rewards"), g("next_states"), g("term"), co.masks_to_bits(masks),
                 float(g("lr")), float(g("gamma")))
    np.testing.assert_array_equal(q, g("q1"))


def _schedule_arrays(lr, eps, steps, n):
    th, lrs = [], []
    for _ in range(steps):
        th.append(orng.explore_threshold(eps.get_value()))
        lrs.append(np.float32(lr.get_value()))
        lr.update(n)
        eps.update(n)
    return np.asarray(th, dtype=np.uint64), np.asarray(lrs, dtype=np.float32)


@pytest.mark.parametrize("name", _cases(TTT))
def test_ttt_run(name):
    g = lambda k: TTT[f"{name}__{k}"]  # noqa: E731
    u = g("u_steps")
    steps, n = u.shape[:2]
    boards, states, masks = co.ttt_reset(g("u_init"))
    q = np.zeros((19683, 9), dtype=np.float32)
    decay = float(g("decay"))
    th, lrs = _schedule_arrays(Exponential(0.1, 1e-5, decay), Exponential(1.0, 0.01, decay), steps, n)
    res = co.run(co.ENV_TTT, q, boards, states, masks, num_states=19683, uniforms=u, steps=steps, eps_thresh=th,
                 lr=lrs, gamma=0.99, record=True)
    assert res["rc"] == 0
    tr = res["trace"]
    for k in ("actions", "rewards", "obs"):
        np.testing.assert_array_equal(tr[k], g(k), err_msg=k)
    np.testing.assert_array_equal(tr["terminated"].astype(bool), g("terminated"))
    er = tr["episode_returns"].ravel()
    np.testing.assert_array_equal(er[~np.isnan(er)], g("history"))
    np.testing.assert_array_equal(q, g("q"))


@pytest.mark.parametrize("name", _cases(MDP))
def test_mdp_run(name):
    g = lambda k: MDP[f"{name}__{k}"]  # noqa: E731
    s, a, n, steps, seed = (int(x) for x in g("cfg"))
    states, masks = co.mdp_reset(g("u_init"), s, a, seed)
    np.testing.assert_array_equal(masks, mdp_mask_bits(states, a, seed))
    q = np.zeros((s, a), dtype=np.float32)
    th, lrs = _schedule_arrays(Exponential(0.5, 1e-3, 0.9999), Exponential(1.0, 0.05, 0.9995), steps, n)
    res = co.run(co.ENV_MDP, q, None, states, masks, num_states=s, env_seed=seed,
                 term_thresh=int(np.ceil(0.05 * 2.0**32)), uniforms=g("u_steps"), steps=steps, eps_thresh=th, lr=lrs,
                 gamma=0.99, empty_all=a > 10, record=True)
    assert res["rc"] == 0
    tr = res["trace"]
    for k in ("actions", "rewards", "obs"):
        np.testing.assert_array_equal(tr[k], g(k), err_msg=k)
    er = tr["episode_returns"].ravel()
    np.testing.assert_array_equal(er[~np.isnan(er)], g("history"))
    np.testing.assert_array_equal(q, g("q"))


def test_counter_stream_equals_predrawn():
    """Generating U on the fly (uniforms=None) must equal feeding the materialised array."""
    s, a, n, steps, seed = 500, 16, 256, 10, 3
    u_init = orng.draw_uniforms(seed, T_INIT, 1, n, 4)[0]
    u = orng.draw_uniforms(seed, 0, steps, n, 4)
    th = np.full(steps, orng.explore_threshold(0.2), dtype=np.uint64)
    lrs = np.full(steps, 0.1, dtype=np.float32)
    out = []
    for uni in (u, None):
        states, masks = co.mdp_reset(u_init, s, a, seed)
        q = np.zeros((s, a), dtype=np.float32)
        co.run(co.ENV_MDP, q, None, states, masks, num_states=s, env_seed=seed, term_thresh=int(np.ceil(0.05 * 2.0**32)),
               uniforms=uni, stream_seed=seed, steps=steps, eps_thresh=th, lr=lrs, gamma=0.99, empty_all=True)
        out.append((q, states))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])
