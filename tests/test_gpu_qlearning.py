"""Unit tests of the engine-backed ``OptimalQLearningBase`` -- the cases of the reference's
``tests/dist_classicrl/algorithms/base_algorithms/test_q_learning_optimal.py`` (T-QLO), run against the
CUDA engine.  Golden values are the reference's (T-QLO:148-282 learn, :337-633 select); the only difference is
the table dtype: float32 here, so literals that are not exactly representable are compared as float32.
"""
from unittest.mock import patch

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def QL():
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase

    return OptimalQLearningBase


def test_init_q_table_shape_and_zeros(QL):
    ql = QL(state_size=3, action_size=4, discount_factor=0.9, seed=123)
    assert ql.q_table.shape == (3, 4)
    assert np.all(ql.q_table == 0)
    assert ql.q_table.dtype == np.float32


def test_accepts_numpy_integer_sizes(QL):
    ql = QL(state_size=np.int32(2), action_size=np.int64(3))
    assert ql.q_table.shape == (2, 3)


def test_set_get_add_q_value(QL):
    ql = QL(3, 4)
    ql.set_q_value(1, 2, 0.5)
    assert ql.get_q_value(1, 2) == 0.5
    ql.set_q_value(0, 0, 1.0)
    ql.add_q_value(0, 0, 0.25)
    assert ql.get_q_value(0, 0) == 1.25


def test_add_q_values_accumulates_duplicates(QL):
    ql = QL(4, 5)
    ql.add_q_values(np.array([2, 2, 3], dtype=np.int32), np.array([4, 4, 1], dtype=np.int32),
                    np.array([0.5, 0.25, 1.0], dtype=np.float64))
    assert ql.get_q_value(2, 4) == 0.75  # accumulates (T-QLO:64-82)
    assert ql.get_q_value(3, 1) == 1.0
    expected = np.zeros((4, 5), dtype=np.float32)
    expected[2, 4], expected[3, 1] = 0.75, 1.0
    assert np.array_equal(ql.q_table, expected)


def test_row_and_column_getters(QL):
    ql = QL(3, 4)
    ql.q_table = np.arange(12, dtype=np.float64).reshape(3, 4)
    assert np.array_equal(ql.get_state_q_values(1), [4, 5, 6, 7])
    assert np.array_equal(ql.get_states_q_values(np.array([2, 0], dtype=np.int32)), [[8, 9, 10, 11], [0, 1, 2, 3]])
    assert np.array_equal(ql.get_action_q_values(2), [2, 6, 10])
    assert np.array_equal(ql.get_actions_q_values(np.array([3, 1], dtype=np.int32)), [[3, 1], [7, 5], [11, 9]])


def test_table_roundtrip_through_device_and_save(QL, tmp_path):
    ql = QL(2, 3)
    ql.q_table = np.array([[1.0, 2.0, 3.0], [4.5, 5.5, 6.5]], dtype=np.float64)
    # a device operation in between forces upload + (after learn) download
    ql.learn(np.array([0]), np.array([0]), np.array([0.0], dtype=np.float32), np.array([1]), np.array([True]), 0.0)
    out = tmp_path / "q_table.npy"
    ql.save(str(out))
    assert np.array_equal(np.load(out), [[1.0, 2.0, 3.0], [4.5, 5.5, 6.5]])


def _invoke_learn(ql, method_name, states, actions, rewards, next_states, terminated, *, lr=1.0, next_action_masks=None):
    method = getattr(ql, method_name)
    if next_action_masks is None:
        method(states, actions, rewards, next_states, terminated, lr)
    else:
        method(states, actions, rewards, next_states, terminated, lr, next_action_masks)


@pytest.mark.parametrize("method_name", ["learn", "learn_iter", "learn_vec"])
def test_learn_single_without_mask(QL, method_name):
    ql = QL(state_size=4, action_size=3, discount_factor=0.5)
    ql.q_table[1] = np.array([1.0, 2.0, 0.5])
    _invoke_learn(ql, method_name, np.array([0], dtype=np.int32), np.array([2], dtype=np.int32),
                  np.array([1.0], dtype=np.float32), np.array([1], dtype=np.int32), np.array([False]))
    assert ql.get_q_value(0, 2) == 2.0  # 1 + 0.5 * 2.0


@pytest.mark.parametrize("method_name", ["learn", "learn_iter", "learn_vec"])
def test_learn_single_with_mask_forces_suboptimal(QL, method_name):
    ql = QL(state_size=4, action_size=3, discount_factor=0.5)
    ql.q_table[2] = np.array([1.0, 3.0, 2.5])
    _invoke_learn(ql, method_name, np.array([0], dtype=np.int32), np.array([0], dtype=np.int32),
                  np.array([0.0], dtype=np.float32), np.array([2], dtype=np.int32), np.array([False]),
                  next_action_masks=np.array([[1, 0, 1]], dtype=np.int32))
    assert ql.get_q_value(0, 0) == 1.25


@pytest.mark.parametrize("method_name", ["learn", "learn_iter", "learn_vec"])
@pytest.mark.parametrize("masked", [False, True])
def test_learn_multiple_with_duplicate_update(QL, method_name, masked):
    ql = QL(state_size=5, action_size=3, discount_factor=0.5)
    ql.q_table[2] = np.array([0.5, 1.5, 1.0])
    ql.q_table[3] = np.array([1.0, 3.0, 2.5])
    masks = np.array([[1, 0, 1]] * 3, dtype=np.int32) if masked else None
    _invoke_learn(ql, method_name, np.array([0, 1, 1], dtype=np.int32), np.array([2, 0, 0], dtype=np.int32),
                  np.array([1.0, 0.0, 2.0], dtype=np.float32), np.array([2, 3, 3], dtype=np.int32),
                  np.array([False, False, False]), next_action_masks=masks)
    if not masked:  # T-QLO:204-232
        assert ql.get_q_value(0, 2) == 1.75
        assert ql.get_q_value(1, 0) == (5.0 if method_name == "learn_vec" else 3.5)
    else:  # T-QLO:235-282
        assert ql.get_q_value(0, 2) == 1.5
        assert ql.get_q_value(1, 0) == (4.5 if method_name == "learn_vec" else 3.25)


def test_learn_terminated_zeroes_bootstrap_and_lr_scales(QL):
    ql = QL(state_size=3, action_size=2, discount_factor=0.9)
    ql.q_table[1] = np.array([10.0, 20.0])
    ql.q_table[0, 1] = 4.0
    ql.learn(np.array([0]), np.array([1]), np.array([2.0], dtype=np.float32), np.array([1]), np.array([True]), 0.5)
    assert ql.get_q_value(0, 1) == 3.0  # 4 + 0.5 * (2 - 4)


def _single(ql, method_name, state, *, exploration_rate, deterministic, action_mask=None):
    if method_name in ("choose_action", "choose_action_vec"):
        return getattr(ql, method_name)(state, exploration_rate, deterministic=deterministic)
    return getattr(ql, method_name)(state, action_mask, exploration_rate, deterministic=deterministic)


def _multi(ql, method_name, states, *, exploration_rate, deterministic, action_masks=None):
    if method_name in ("choose_actions", "choose_actions_iter", "choose_actions_vec_iter"):
        return getattr(ql, method_name)(states, exploration_rate, deterministic=deterministic, action_masks=action_masks)
    if method_name == "choose_actions_vec":
        return ql.choose_actions_vec(states, exploration_rate, deterministic=deterministic)
    return ql.choose_masked_actions_vec(states, action_masks, exploration_rate, deterministic=deterministic)


@pytest.mark.parametrize("method_name", ["choose_action", "choose_action_vec"])
def test_single_deterministic_unmasked_unique_max(QL, method_name):
    ql = QL(state_size=2, action_size=3, discount_factor=0.9, seed=123)
    ql.q_table[0] = np.array([0.1, 0.8, 0.2])
    assert _single(ql, method_name, 0, exploration_rate=0.0, deterministic=True) == 1


@pytest.mark.parametrize("method_name", ["choose_masked_action", "choose_masked_action_vec"])
def test_single_deterministic_masked_unique_max(QL, method_name):
    ql = QL(state_size=2, action_size=3, discount_factor=0.9, seed=123)
    ql.q_table[0] = np.array([0.3, 1.0, 0.7])
    assert _single(ql, method_name, 0, exploration_rate=0.0, deterministic=True, action_mask=[1, 0, 1]) == 2


@pytest.mark.parametrize("method_name", ["choose_action", "choose_action_vec"])
def test_single_nondeterministic_unmasked_explore(QL, method_name):
    ql = QL(state_size=2, action_size=3, discount_factor=0.9, seed=123)
    ql.q_table[0] = np.array([0.1, 0.8, 0.2])
    draw = "uniform" if method_name == "choose_action" else "random"
    with patch.object(ql._rng, draw, return_value=0.0), patch.object(ql._rng, "randint", return_value=2):
        assert _single(ql, method_name, 0, exploration_rate=1.0, deterministic=False) == 2


@pytest.mark.parametrize("method_name", ["choose_masked_action", "choose_masked_action_vec"])
def test_single_nondeterministic_masked_single_option(QL, method_name):
    ql = QL(state_size=2, action_size=3, discount_factor=0.9, seed=123)
    ql.q_table[0] = np.array([0.3, 1.0, 0.7])
    draw = "uniform" if method_name == "choose_masked_action" else "random"
    with patch.object(ql._rng, draw, return_value=0.0), patch.object(ql._rng, "choice", return_value=1):
        assert _single(ql, method_name, 0, exploration_rate=1.0, deterministic=False, action_mask=[0, 1, 0]) == 1
    # and without mocks: only action 1 is legal, whatever the stream says
    for _ in range(5):
        assert _single(ql, method_name, 0, exploration_rate=1.0, deterministic=False, action_mask=[0, 1, 0]) == 1


@pytest.mark.parametrize("method_name", ["choose_actions", "choose_actions_iter", "choose_actions_vec_iter", "choose_actions_vec"])
def test_multi_deterministic_unmasked_unique_max(QL, method_name):
    ql = QL(state_size=4, action_size=3, discount_factor=0.9, seed=123)
    ql.q_table[0] = np.array([0.2, 0.9, 0.1])
    ql.q_table[2] = np.array([0.5, 0.4, 0.7])
    actions = _multi(ql, method_name, np.array([0, 2], dtype=np.int32), exploration_rate=0.0, deterministic=True)
    assert actions.dtype == np.int32
    assert np.array_equal(actions, [1, 2])


@pytest.mark.parametrize("method_name", ["choose_actions", "choose_actions_iter", "choose_actions_vec_iter", "choose_masked_actions_vec"])
def test_multi_deterministic_masked_unique_max(QL, method_name):
    ql = QL(state_size=4, action_size=3, discount_factor=0.9, seed=123)
    ql.q_table[0] = np.array([0.2, 0.9, 0.1])
    ql.q_table[2] = np.array([0.5, 0.4, 0.7])
    masks = np.array([[1, 0, 1], [1, 1, 0]], dtype=np.int32)
    actions = _multi(ql, method_name, np.array([0, 2], dtype=np.int32), exploration_rate=0.0, deterministic=True, action_masks=masks)
    assert np.array_equal(actions, [0, 0])


@pytest.mark.parametrize("method_name", ["choose_actions", "choose_actions_iter", "choose_actions_vec_iter", "choose_actions_vec"])
def test_multi_nondeterministic_unmasked_explore(QL, method_name):
    ql = QL(state_size=4, action_size=3, discount_factor=0.9, seed=123)
    states = np.array([0, 1, 2], dtype=np.int32)
    if method_name == "choose_actions_vec":
        with patch.object(ql, "_np_rng") as np_rng_mock:
            np_rng_mock.random.return_value = np.array([0.0, 0.0, 0.0])
            np_rng_mock.integers.return_value = np.array([0, 1, 2])
            actions = _multi(ql, method_name, states, exploration_rate=1.0, deterministic=False)
    else:
        draw = "uniform" if method_name in ("choose_actions", "choose_actions_iter") else "random"
        with patch.object(ql._rng, draw, return_value=0.0), patch.object(ql._rng, "randint", side_effect=[0, 1, 2]):
            actions = _multi(ql, method_name, states, exploration_rate=1.0, deterministic=False)
    assert np.array_equal(actions, [0, 1, 2])


@pytest.mark.parametrize("method_name", ["choose_actions", "choose_actions_iter", "choose_actions_vec_iter", "choose_masked_actions_vec"])
def test_multi_nondeterministic_masked_explore(QL, method_name):
    ql = QL(state_size=4, action_size=4, discount_factor=0.9, seed=123)
    states = np.array([0, 1, 2], dtype=np.int32)
    masks = np.array([[1, 0, 0, 0], [0, 1, 1, 0], [0, 0, 0, 1]], dtype=np.int32)
    if method_name == "choose_masked_actions_vec":
        with patch.object(ql, "_np_rng") as np_rng_mock, patch.object(ql._rng, "choice", side_effect=[0, 2, 3]):
            np_rng_mock.random.return_value = np.array([0.0, 0.0, 0.0])
            actions = _multi(ql, method_name, states, exploration_rate=1.0, deterministic=False, action_masks=masks)
    else:
        draw = "uniform" if method_name in ("choose_actions", "choose_actions_iter") else "random"
        with patch.object(ql._rng, draw, return_value=0.0), patch.object(ql._rng, "choice", side_effect=[0, 2, 3]):
            actions = _multi(ql, method_name, states, exploration_rate=1.0, deterministic=False, action_masks=masks)
    assert np.array_equal(actions, [0, 2, 3])
    # engine stream, no mocks: every pick is legal
    for _ in range(10):
        a = _multi(ql, method_name, states, exploration_rate=1.0, deterministic=False, action_masks=masks)
        assert a[0] == 0 and a[1] in (1, 2) and a[2] == 3


@pytest.mark.parametrize("method_name", ["choose_action", "choose_action_vec"])
def test_tie_breaking_goes_through_rng_choice(QL, method_name):
    ql = QL(state_size=1, action_size=3, discount_factor=0.9, seed=123)
    ql.q_table[0] = np.array([0.9, 0.9, 0.1])
    with patch.object(ql._rng, "choice", side_effect=[0, 1]) as ch:
        a1 = _single(ql, method_name, 0, exploration_rate=0.0, deterministic=True)
        a2 = _single(ql, method_name, 0, exploration_rate=0.0, deterministic=True)
        assert [int(x) for x in ch.call_args_list[0].args[0]] == [0, 1]  # the tie set, ascending
    assert (a1, a2) == (0, 1)


@pytest.mark.parametrize("method_name", ["choose_actions", "choose_actions_iter", "choose_actions_vec_iter", "choose_masked_actions_vec"])
def test_tie_breaking_masked_array(QL, method_name):
    ql = QL(state_size=3, action_size=3, discount_factor=0.9, seed=123)
    ql.q_table[0] = np.array([0.7, 0.7, 0.7])
    ql.q_table[1] = np.array([0.1, 0.9, 0.9])
    masks = np.array([[1, 0, 1], [0, 1, 1]], dtype=np.int32)
    with patch.object(ql._rng, "choice", side_effect=[2, 1]):
        actions = _multi(ql, method_name, np.array([0, 1], dtype=np.int32), exploration_rate=0.0, deterministic=True,
                         action_masks=masks)
    assert np.array_equal(actions, [2, 1])


def test_engine_stream_tie_breaking_is_uniform_over_tie_set(QL):
    """Without mocks the engine's own stream breaks ties: only tied legal actions, each of them reachable."""
    ql = QL(state_size=2, action_size=9, discount_factor=0.9, seed=7)
    ql.q_table[0] = np.array([0.5, 0.1, 0.5, 0.5, 0.0, 0.5, 0.9, 0.5, 0.2])
    masks = np.tile(np.array([[1, 1, 1, 0, 1, 1, 0, 1, 1]], dtype=np.int8), (4096, 1))
    a = ql.choose_actions(np.zeros(4096, dtype=np.int64), 0.0, action_masks=masks)
    assert set(np.unique(a)) == {0, 2, 5, 7}
    counts = np.bincount(a, minlength=9)[[0, 2, 5, 7]]
    assert counts.min() > 800  # ~1024 each


def test_empty_mask_behaviour_follows_the_dispatcher(QL):
    ql = QL(state_size=2, action_size=3, seed=1)
    assert ql.choose_actions(np.array([0, 1]), 0.5, action_masks=np.zeros((2, 3), dtype=np.int32)).tolist() == [-1, -1]
    ql16 = QL(state_size=2, action_size=16, seed=1)
    with pytest.raises(IndexError):  # vec variant explores on an empty candidate array (QLO:465,470)
        ql16.choose_actions(np.array([0, 1]), 1.0, action_masks=np.zeros((2, 16), dtype=np.int32))
    a = ql16.choose_actions(np.array([0, 1]), 0.0, action_masks=np.zeros((2, 16), dtype=np.int32))
    assert ((a >= 0) & (a < 16)).all()  # exploit: every action ties at -inf (QLO:467-470)


def test_mask_shape_is_checked(QL):
    ql = QL(state_size=2, action_size=3)
    with pytest.raises(AssertionError):
        ql.choose_actions(np.array([0, 1]), 0.0, action_masks=np.ones((2, 4), dtype=np.int32))


def test_device_tensor_inputs_match_host_inputs(QL):
    rng = np.random.default_rng(0)
    s, a, n = 500, 16, 2048
    q0 = rng.standard_normal((s, a)).astype(np.float32)
    st, ac = rng.integers(0, s, n), rng.integers(0, a, n)
    rw, s2, tm = rng.standard_normal(n).astype(np.float32), rng.integers(0, s, n), rng.random(n) < 0.1
    masks = (rng.random((n, a)) < 0.5).astype(np.int8)
    masks[:, 0] = 1
    host, dev = QL(s, a, 0.9, seed=3), QL(s, a, 0.9, seed=3)
    host.q_table, dev.q_table = q0.copy(), q0.copy()
    ah = host.choose_actions(st, 0.3, action_masks=masks)
    ad = dev.choose_actions(torch.from_numpy(st).cuda(), 0.3, action_masks=torch.from_numpy(masks).cuda())
    assert np.array_equal(ah, ad.cpu().numpy())
    host.learn(st, ac, rw, s2, tm, 0.2, masks)
    dev.learn(*(torch.from_numpy(x).cuda() for x in (st, ac, rw, s2, tm)), 0.2, torch.from_numpy(masks).cuda())
    assert np.array_equal(host.q_table, dev.q_table)
