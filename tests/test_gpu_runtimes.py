"""Runtime-level parity on the GPU: the cases of the reference's runtime tests (T-RT, T-EV) plus full
trajectories through ``SingleThreadQLearning.run_steps`` against fixtures produced by the REAL reference
(tests/golden, oracle/make_golden.py).  Fused and unfused paths must agree bit for bit."""
import os
from collections.abc import Sequence

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TTT = np.load(os.path.join(GOLDEN, "ttt_traj.npz"))
MDP = np.load(os.path.join(GOLDEN, "mdp_traj.npz"))
BANDIT = np.load(os.path.join(GOLDEN, "bandit_traj.npz"))


def _cases(npz):
    return sorted({k.split("__")[0] for k in npz.files})


@pytest.fixture(scope="module")
def api():
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    import types

    from dist_classicrl_b200 import environments, rng, schedules
    from dist_classicrl_b200.algorithms import runtime
    from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase

    return types.SimpleNamespace(QL=OptimalQLearningBase, rt=runtime, env=environments, sch=schedules, rng=rng)


class DeterministicRNG:
    """The shim of the reference's runtime tests (T-RT:17-45): always explore, always pick action 1."""

    def uniform(self, _a=0.0, _b=1.0):
        return 0.0

    def random(self):
        return 0.0

    def randint(self, _a, _b):
        return 1

    def choice(self, seq: Sequence[int] | np.ndarray):
        arr = np.asarray(seq)
        return 1 if (arr == 1).any() else int(arr[0])


def _bandit_runtime(api, runtime_cls):
    algo = api.QL(state_size=1, action_size=2, discount_factor=1.0, seed=0)
    algo._rng = DeterministicRNG()
    rt = runtime_cls(algorithm=algo, lr_schedule=api.sch.ConstantSchedule(1.0), exploration_rate_schedule=api.sch.LinearSchedule(1.0, 1.0))
    return algo, rt


def test_single_thread_run_steps_random_actions_and_updates(api):  # T-RT:77-98
    algo, rt = _bandit_runtime(api, api.rt.SingleThreadQLearning)
    env = api.env.make_bandit_vec_env(1, episode_len=5)
    avg, history, _env, state_dict = rt.run_steps(steps=5, env=env, curr_state_dict=None)
    assert history == [5.0] and avg == 5.0
    assert algo.q_table.shape == (1, 2)
    assert algo.q_table[0, 1] == 1.0
    assert rt.lr_schedule.get_value() == 1.0
    assert rt.exploration_rate_schedule.get_value() == 6.0
    assert isinstance(state_dict["states"], np.ndarray)


def test_parallel_run_steps_two_envs_10_steps(api):  # T-RT:132-160
    _algo, rt = _bandit_runtime(api, api.rt.ParallelQLearning)
    envs = [api.env.make_bandit_vec_env(1, episode_len=5) for _ in range(2)]
    try:
        rt.init_training()
        avg, history, _envs, states_list = rt.run_steps(steps=10, env=envs, curr_state_dict=None)
        assert history == [5.0, 5.0] and avg == 5.0
        assert rt.algorithm.q_table[0, 1] == 1.0
        assert rt.lr_schedule.get_value() == 1.0
        assert rt.exploration_rate_schedule.get_value() == 11.0
        assert len(states_list) == 2 and all("states" in d for d in states_list)
    finally:
        rt.close_training()


@pytest.mark.parametrize("n_envs,steps,episodes", [(1, 10, 3), (3, 30, 6), (4, 40, 8)])
def test_evaluate_steps_and_episodes(api, n_envs, steps, episodes):  # T-EV:57-116
    algo = api.QL(state_size=1, action_size=2, discount_factor=0.99, seed=0)
    algo.q_table[0] = np.array([0.0, 1.0])
    rt = api.rt.SingleThreadQLearning(algo, api.sch.ConstantSchedule(0.0), api.sch.ConstantSchedule(0.0))
    total, history = rt.evaluate_steps(api.env.make_bandit_vec_env(n_envs, episode_len=10), steps=steps)
    expected_len = (steps // n_envs // 10) * n_envs
    assert history == [10.0] * expected_len and total == 10.0 * expected_len
    total, history = rt.evaluate_episodes(api.env.make_bandit_vec_env(n_envs, episode_len=10), episodes=episodes)
    assert history == [10.0] * episodes and total == 10.0 * episodes


def _ttt_setup(api, name):
    g = lambda k: TTT[f"{name}__{k}"]  # noqa: E731
    u = g("u_steps")
    n = u.shape[1]
    env = api.env.TicTacToeVecEnv(n)
    algo = api.QL(19683, 9, 0.99, seed=0)
    decay = float(g("decay"))
    rt = api.rt.SingleThreadQLearning(algo, api.sch.ExponentialSchedule(0.1, 1e-5, decay), api.sch.ExponentialSchedule(1.0, 0.01, decay))
    stream = api.rng.PredrawnUniforms(u)
    algo._rng = stream
    env._rng = stream
    env.attach(algo)
    states, infos = env.reset_with(g("u_init"))
    sd = {"states": states, "infos": infos, "rewards": np.zeros(n, dtype=np.float32)}
    return g, u, env, algo, rt, sd


@pytest.mark.parametrize("name", _cases(TTT))
def test_tictactoe_run_steps_fused_matches_reference(api, name):
    g, u, env, algo, rt, sd = _ttt_setup(api, name)
    trace = {}
    avg, history, _env, out = rt.run_steps(u.shape[0], env, sd, trace=trace)
    assert rt._can_fuse(env)
    for k in ("actions", "rewards", "obs"):
        np.testing.assert_array_equal(np.concatenate(trace[k]), g(k), err_msg=k)
    np.testing.assert_array_equal(np.concatenate(trace["terminated"]).astype(bool), g("terminated"))
    np.testing.assert_array_equal(np.asarray(history, dtype=np.float32), g("history"))
    assert avg == pytest.approx(float(np.mean(g("history"))), rel=1e-6)
    np.testing.assert_array_equal(algo.q_table, g("q"))  # bit exact vs the reference (fp32 rewards)
    np.testing.assert_allclose(algo.q_table, g("q_f64r"), rtol=1e-6, atol=1e-9)  # reference as-is (fp64 rewards)
    assert rt.lr_schedule.get_value() == float(g("lr_end"))
    assert rt.exploration_rate_schedule.get_value() == float(g("eps_end"))
    np.testing.assert_array_equal(out["states"]["observation"], g("obs")[-1])
    assert out["states"]["action_mask"].shape == (u.shape[1], 9)


@pytest.mark.parametrize("name", ["n16"])
def test_tictactoe_unfused_equals_fused(api, name):
    """choose_actions -> env.step -> learn with host arrays (the reference's own loop, BRT:184-222) must
    give the same trajectory and table as the fused kernel."""
    g, u, env, algo, rt, sd = _ttt_setup(api, name)
    algo._rng = api.rng.PredrawnUniforms(u)  # the reference keeps two independent generators
    env._rng = api.rng.PredrawnUniforms(u)
    rt._can_fuse = lambda _env: False  # force the reference-shaped loop
    avg, history, _env, out = rt.run_steps(u.shape[0], env, sd)
    np.testing.assert_array_equal(np.asarray(history, dtype=np.float32), g("history"))
    np.testing.assert_array_equal(algo.q_table, g("q"))
    np.testing.assert_array_equal(out["states"]["observation"], g("obs")[-1])


@pytest.mark.parametrize("name", _cases(MDP))
def test_hash_mdp_run_steps_matches_reference(api, name):
    g = lambda k: MDP[f"{name}__{k}"]  # noqa: E731
    s, a, n, steps, seed = (int(x) for x in g("cfg"))
    env = api.env.HashMDPVecEnv(n, s, a, env_seed=seed, p_term=0.05)
    algo = api.QL(s, a, 0.99, seed=0)
    rt = api.rt.SingleThreadQLearning(algo, api.sch.ExponentialSchedule(0.5, 1e-3, 0.9999), api.sch.ExponentialSchedule(1.0, 0.05, 0.9995))
    stream = api.rng.PredrawnUniforms(g("u_steps"))
    algo._rng = stream
    env._rng = stream
    env.attach(algo)
    states, infos = env.reset_with(g("u_init"))
    trace = {}
    _avg, history, _env, out = rt.run_steps(steps, env, {"states": states, "infos": infos, "rewards": np.zeros(n, dtype=np.float32)},
                                            trace=trace)
    for k in ("actions", "rewards", "obs"):
        np.testing.assert_array_equal(np.concatenate(trace[k]), g(k), err_msg=k)
    np.testing.assert_array_equal(np.asarray(history, dtype=np.float32), g("history"))
    np.testing.assert_array_equal(algo.q_table, g("q"))
    assert rt.lr_schedule.get_value() == float(g("lr_end"))


def test_bandit_fused_matches_reference(api):
    u = BANDIT["u_steps"]
    n = u.shape[1]
    env = api.env.make_bandit_vec_env(n, episode_len=5)
    algo = api.QL(1, 2, 0.9, seed=0)
    rt = api.rt.SingleThreadQLearning(algo, api.sch.ConstantSchedule(0.25), api.sch.LinearSchedule(0.9, -0.01))
    stream = api.rng.PredrawnUniforms(u)
    algo._rng = stream
    env._rng = stream
    trace = {}
    _avg, history, _env, _sd = rt.run_steps(u.shape[0], env, None, trace=trace)
    for k in ("actions", "rewards", "obs"):
        np.testing.assert_array_equal(np.concatenate(trace[k]), BANDIT[k], err_msg=k)
    np.testing.assert_array_equal(np.asarray(history, dtype=np.float32), BANDIT["history"])
    np.testing.assert_array_equal(algo.q_table, BANDIT["q"])
    assert rt.exploration_rate_schedule.get_value() == float(BANDIT["eps_end"])


def test_counter_stream_fused_equals_unfused_mdp(api):
    """Default CounterRNG streams: on-device generation in the fused kernel == per-call generation."""
    s, a, n, steps = 2000, 16, 1024, 25
    out = []
    for fused in (True, False):
        env = api.env.HashMDPVecEnv(n, s, a, env_seed=4, seed=11)
        algo = api.QL(s, a, 0.95, seed=9)
        env.attach(algo)
        rt = api.rt.SingleThreadQLearning(algo, api.sch.ConstantSchedule(0.3), api.sch.ExponentialSchedule(1.0, 0.1, 0.9999))
        if not fused:
            rt._can_fuse = lambda _env: False
        _avg, history, _env, sd = rt.run_steps(steps, env, None)
        out.append((np.asarray(history, dtype=np.float32), algo.q_table.copy(), sd["states"]["observation"].copy()))
    for x, y in zip(out[0], out[1]):
        np.testing.assert_array_equal(x, y)


def test_train_entry_point_with_validation(api):
    """train(): chunks of val_every_n_steps, one evaluation per chunk (BRT:156-173), TPB-style schedules."""
    n = 64
    env = api.env.TicTacToeVecEnv(n, seed=1)
    val_env = api.env.TicTacToeVecEnv(1, seed=2)
    algo = api.QL(env.single_observation_space["observation"].n, env.single_action_space.n, 0.99, seed=0)
    env.attach(algo)
    val_env.attach(algo)
    rt = api.rt.SingleThreadQLearning(algo, api.sch.ExponentialSchedule(0.1, 1e-5, 0.995), api.sch.ExponentialSchedule(1.0, 0.01, 0.995))
    history, val_history, _env, sd = rt.train(env, steps=200, val_env=val_env, val_every_n_steps=100, val_episodes=10)
    assert len(val_history) == 2
    assert len(history) > 100 and set(np.unique(history)) <= {-1.0, 0.0, 1.0}
    assert np.abs(algo.q_table).max() > 0
    assert set(sd) == {"states", "infos", "rewards", "episode_rewards"}
    with pytest.raises(AssertionError):
        rt.train(env, steps=10, val_env=val_env, val_every_n_steps=10)


def test_state_dict_returns_are_pinned_in_place_and_match_the_oracle(api):
    """A state dictionary's running returns (plain NumPy, STR:57) large enough to be page-locked in place: the array
    object is updated in place across resumed ``run_steps`` calls and equals the oracle's accumulator; the episode
    statistics of summary mode equal the oracle's too."""
    import math

    from dist_classicrl_b200 import hostmem
    from oracle import c_oracle as co
    from oracle import rng as orng

    s, a, n, calls, k = 5000, 16, 32768, 4, 3
    u = orng.draw_uniforms(3, 0, calls * k, n, 4)
    u0 = orng.draw_uniforms(3, 0xFFFFFFFF, 1, n, 4)[0]
    env = api.env.HashMDPVecEnv(n, s, a, env_seed=7, p_term=0.05)
    algo = api.QL(s, a, 0.9, seed=0)
    env.attach(algo)
    env.reset_with(u0)
    algo._rng = env._rng = api.rng.PredrawnUniforms(u)
    rt = api.rt.SingleThreadQLearning(algo, api.sch.ConstantSchedule(0.25), api.sch.ConstantSchedule(0.2))
    rt.history_mode = "summary"
    rewards = np.zeros(n, dtype=np.float32)
    sd = {"states": None, "infos": {}, "rewards": rewards}
    ep_sum, ep_cnt = 0.0, 0
    for _ in range(calls):
        _mean, _hist, _env, sd = rt.run_steps(k, env, sd)
        assert sd["rewards"] is rewards
        ep_sum += rt.last_episode_sum
        ep_cnt += rt.last_episode_count
    assert hostmem._registered.get(rewards.ctypes.data) == rewards.nbytes
    assert torch.from_numpy(rewards).is_pinned()
    # oracle on the same streams
    states, masks = co.mdp_reset(u0, s, a, 7)
    q = np.zeros((s, a), dtype=np.float32)
    res = co.run(co.ENV_MDP, q, None, states, masks, num_states=s, env_seed=7, term_thresh=int(math.ceil(0.05 * 2.0**32)),
                 uniforms=u, steps=calls * k, eps_thresh=np.full(calls * k, orng.explore_threshold(0.2), dtype=np.uint64),
                 lr=np.full(calls * k, 0.25, dtype=np.float32), gamma=0.9, empty_all=True)
    assert res["rc"] == 0
    np.testing.assert_array_equal(rewards, res["agent_rewards"])
    np.testing.assert_array_equal(algo.q_table, q)
    assert ep_cnt == res["ep_count"] and abs(ep_sum - res["ep_sum"]) <= 1e-6 * max(1.0, abs(res["ep_sum"]))
    ptr = rewards.ctypes.data
    del rewards, sd
    import gc

    gc.collect()
    assert ptr not in hostmem._registered


def _async_rank(api, msg, algo, steps, episode_len, val_every, val_steps, batch, out):
    from dist_classicrl_b200.environments.rigged_two_armed_bandit import make_bandit_vec_env

    rt = api.rt.DistAsyncQLearning(algo, api.sch.ConstantSchedule(1.0), api.sch.LinearSchedule(1.0, 1.0), messenger=msg)
    out[msg.rank] = (rt, rt.train(env=make_bandit_vec_env(1, episode_len), steps=steps, val_env=make_bandit_vec_env(1, episode_len),
                                  val_every_n_steps=val_every, val_steps=val_steps, val_episodes=None, curr_state_dict={}, batch_size=batch))


@pytest.mark.parametrize("world", [2, 3])
def test_dist_async_train_skips_validation_and_updates_q_table(api, world):  # T-MPI:100-147 (ranks as threads)
    import threading

    from dist_classicrl_b200.algorithms.runtime.q_learning_async_dist import ThreadMessenger

    algos = []
    for _ in range(world):
        algo = api.QL(state_size=1, action_size=2, discount_factor=1.0, seed=0)
        algo._rng = DeterministicRNG()
        algos.append(algo)
    out = {}
    ts = [threading.Thread(target=_async_rank, args=(api, m, a, 5, 6, 6, 6, 8, out)) for m, a in zip(ThreadMessenger.group(world), algos)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert len(out) == world
    rt, (hist, val, envs, state) = out[0]
    assert val == [] and isinstance(hist, list) and envs is None and state is None
    assert rt.algorithm.q_table.shape == (1, 2) and rt.algorithm.q_table[0, 1] == 5.0
    assert rt.lr_schedule.get_value() == 1.0 and rt.exploration_rate_schedule.get_value() == 6.0
    for r in range(1, world):
        _rt, (hist, val, env, state) = out[r]
        assert hist == [] and val == [] and env is not None and "states" in state


def test_dist_async_train_with_validation(api):  # T-MPI:150-205
    import threading

    from dist_classicrl_b200.algorithms.runtime.q_learning_async_dist import ThreadMessenger

    out = {}
    algos = []
    for _ in range(2):
        algo = api.QL(state_size=1, action_size=2, discount_factor=1.0, seed=0)
        algo._rng = DeterministicRNG()
        algos.append(algo)
    ts = [threading.Thread(target=_async_rank, args=(api, m, a, 6, 5, 3, 5, 2, out)) for m, a in zip(ThreadMessenger.group(2), algos)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    rt, (hist, val, _envs, _state) = out[0]
    assert len(val) == 2 and all(v == 5.0 for v in val)
    assert [float(h) for h in hist] == [5.0]
    assert rt.exploration_rate_schedule.get_value() == 7.0


@pytest.mark.parametrize("s,n", [(3000, 4096), (200000, 2048)])
def test_fused_accumulate_mode_matches_learn_vec_step_by_step(api, s, n):
    """``td_update = "accumulate"``: the fused loop with the plain-atomics update = select -> env step -> ``learn_vec``
    (QLO:819-891).  Checked one vector step at a time against the oracle started from the engine's own table of the
    step before: actions, rewards, flags and next states bit-exact; the table within 1e-6 (the reference evaluates
    ``gamma * max * (1 - terminated)`` in float64 because of the int64 factor and rounds on the scatter, the engine
    works in fp32; the increments of one cell are summed by atomics in any order, np.add.at sums them in agent order)."""
    from oracle import qlearning as oq
    from oracle import rng as orng
    from oracle.envs import HashMDPVec

    a, steps, gamma, lr, eps = 16, 5, 0.9, 0.25, 0.3
    u = orng.draw_uniforms(5, 0, steps, n, 4)
    u0 = orng.draw_uniforms(5, 0xFFFFFFFF, 1, n, 4)[0]
    env = api.env.HashMDPVecEnv(n, s, a, env_seed=3, p_term=0.05)
    algo = api.QL(s, a, gamma, seed=0)
    algo.q_table = np.random.default_rng(2).random((s, a), dtype=np.float32)
    env.attach(algo)
    states, infos = env.reset_with(u0)
    algo._rng = env._rng = api.rng.PredrawnUniforms(u)
    rt = api.rt.SingleThreadQLearning(algo, api.sch.ConstantSchedule(lr), api.sch.ConstantSchedule(eps))
    rt.td_update = "accumulate"
    oenv = HashMDPVec(n, s, a, seed=3, p_term=0.05)
    ostates, _ = oenv.reset(u0)
    sd = {"states": states, "infos": infos, "rewards": np.zeros(n, dtype=np.float32)}
    for t in range(steps):
        q_before = algo.q_table.copy()
        trace = {}
        _m, _h, _e, sd = rt.run_steps(1, env, sd, trace=trace)
        act = oq.select(q_before, ostates["observation"], ostates["action_mask"], eps, u[t])
        nxt, rew, term, _trunc, _ = oenv.step(act, u[t])
        np.testing.assert_array_equal(trace["actions"][0][0], act)
        np.testing.assert_array_equal(trace["rewards"][0][0], rew)
        np.testing.assert_array_equal(trace["terminated"][0][0].astype(bool), term)
        np.testing.assert_array_equal(trace["obs"][0][0], nxt["observation"])
        q_ref = q_before.copy()
        oq.learn_accumulate(q_ref, ostates["observation"], act, rew, nxt["observation"], term, lr, gamma, nxt["action_mask"])
        q_gpu = algo.q_table
        np.testing.assert_allclose(q_gpu, q_ref, rtol=1e-6, atol=1e-6)
        cells = ostates["observation"].astype(np.int64) * a + act
        untouched = np.bincount(cells, minlength=s * a).reshape(s, a) == 0
        np.testing.assert_array_equal(q_gpu[untouched], q_before[untouched])
        assert not np.array_equal(q_gpu, q_before)
        algo.q_table = q_ref  # both sides continue from the reference's table
        ostates = nxt
