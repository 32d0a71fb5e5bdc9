"""TicTacToe environment tests on the GPU -- the properties pinned by the reference's
``tests/dist_classicrl/environment/test_tiktaktoe_mod.py`` (T-TTT), exercised through the device vector env
(boards are set directly, like the reference tests assign ``env.board``)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Env():
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    from dist_classicrl_b200.environments import TicTacToeVecEnv

    return TicTacToeVecEnv


def _step(env, boards, marks, actions):
    env.set_boards(np.asarray(boards), np.asarray(marks))
    return env.step(np.asarray(actions))


def test_reset_shapes_and_initial_position(Env):  # T-TTT:267-275
    env = Env(512, seed=3)
    obs, info = env.reset()
    assert obs["observation"].shape == (512,) and obs["action_mask"].shape == (512, 9)
    assert obs["observation"].dtype == np.int64 and obs["action_mask"].dtype == np.int64
    b, marks = env.boards, env.agent_marks
    n_marks = (b != 0).sum(axis=1)
    assert set(np.unique(n_marks)) == {0, 1}  # empty board (agent starts) or one machine opening
    assert ((marks == 1) == (n_marks == 0)).all()  # agent is mark 1 iff it starts (TTT:99-107)
    assert (b[n_marks == 1].max(axis=1) == 1).all()  # the opening machine plays mark 1
    np.testing.assert_array_equal(obs["action_mask"], (b == 0).astype(np.int64))
    np.testing.assert_array_equal(obs["observation"], b.astype(np.int64) @ (3 ** np.arange(8, -1, -1)))
    assert 0.4 < (n_marks == 0).mean() < 0.6


def test_agent_win_loss_draw_rewards(Env):  # T-TTT:150-264
    env = Env(4, seed=0)
    env.reset()
    boards = [
        [1, 1, 0, 2, 2, 0, 0, 0, 0],  # agent (1) completes the top row -> +1
        [1, 1, 2, 2, 2, 1, 0, 1, 0],  # agent plays 8 (no line); the machine's only cell 6 completes 2-4-6 -> -1
        [1, 2, 1, 1, 2, 2, 2, 1, 0],  # agent fills the last cell -> draw, 0
        [2, 2, 0, 1, 1, 0, 0, 0, 0],  # agent is mark 2 here and completes the top row -> +1
    ]
    marks = [1, 1, 1, 2]
    obs, r, term, trunc, _ = _step(env, boards, marks, [2, 8, 8, 2])
    assert r[0] == 1.0 and term[0] and not trunc[0]
    assert r[2] == 0.0 and term[2]
    assert r[3] == 1.0 and term[3]
    assert r[1] == -1.0 and term[1]
    # SAME_STEP autoreset: terminated envs already show a fresh position
    assert ((env.boards != 0).sum(axis=1) <= 1).all()


def test_invalid_move_raises(Env):  # T-TTT:247-264 / TTT:130
    env = Env(2, seed=0)
    env.reset()
    with pytest.raises(AssertionError, match="Invalid move"):
        _step(env, [[1, 0, 0, 0, 0, 0, 0, 0, 0], [0] * 9], [2, 1], [0, 4])


def test_machine_forced_move_and_step_leaves_free_cells(Env):  # T-TTT:121-147, 278-294
    env = Env(2, seed=5)
    env.reset()
    # env 0: after the agent plays 0 the only free cell is 4 and nobody wins -> machine plays 4, board full, draw
    boards = [[0, 2, 1, 2, 0, 1, 1, 1, 2], [0] * 9]
    obs, r, term, _, _ = _step(env, boards, [1, 1], [0, 0])
    assert term[0] and r[0] == 0.0
    # env 1: agent opened at 0, machine answered somewhere -> 7 free cells, agent mark at 0
    b = env.boards[1]
    assert b[0] == 1 and (b == 0).sum() == 7 and (b == 2).sum() == 1
    assert not term[1] and r[1] == 0.0
    assert obs["action_mask"][1].sum() == 7


def test_random_full_games_smoke(Env):  # T-TTT:297-313
    env = Env(256, seed=9)
    obs, _ = env.reset()
    rng = np.random.default_rng(0)
    finished = 0
    for _ in range(60):
        mask = obs["action_mask"]
        actions = np.array([rng.choice(np.nonzero(m)[0]) for m in mask])
        obs, r, term, trunc, _ = env.step(actions)
        assert set(np.unique(r)) <= {-1.0, 0.0, 1.0}
        assert (r[~term] == 0).all() and not trunc.any()
        finished += int(term.sum())
    assert finished > 256 * 5
