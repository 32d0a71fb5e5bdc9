"""device evaluation loops (BRT:293-384), save/load of the table (QLO:252-261) and the GPU
``throughput_benchmark`` command line (TPB:105-323)."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")


def _trained(n=64, steps=40):
    from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase
    from dist_classicrl_b200.algorithms.runtime import SingleThreadQLearning
    from dist_classicrl_b200.environments import TicTacToeVecEnv
    from dist_classicrl_b200.schedules import ExponentialSchedule

    algo = OptimalQLearningBase(19683, 9, 0.99, seed=3)
    env = TicTacToeVecEnv(n, seed=5).attach(algo)
    rt = SingleThreadQLearning(algo, ExponentialSchedule(0.1, 1e-5, 0.995), ExponentialSchedule(1.0, 0.01, 0.995))
    rt.run_steps(steps, env)
    return algo, rt


@pytest.mark.parametrize("mode", ["steps", "episodes"])
def test_device_evaluation_equals_host_loop(mode):
    from dist_classicrl_b200.environments import TicTacToeVecEnv

    algo, rt = _trained()
    before = np.array(algo.q_table, copy=True)
    lr_before, eps_before = rt.lr_schedule.get_value(), rt.exploration_rate_schedule.get_value()
    t_algo = algo._rng.t
    fused_env = TicTacToeVecEnv(16, seed=1).attach(algo)
    res_fused = rt.evaluate_steps(fused_env, 400) if mode == "steps" else rt.evaluate_episodes(fused_env, 25)
    assert np.array_equal(np.asarray(algo.q_table), before), "evaluation must not touch the table"
    assert rt.lr_schedule.get_value() == lr_before and rt.exploration_rate_schedule.get_value() == eps_before
    # the reference-shaped host loop (choose_actions(deterministic=True) -> env.step) on the same streams
    algo._rng.t = t_algo
    host_env = TicTacToeVecEnv(16, seed=1, output="torch-unfused").attach(algo)
    res_host = rt.evaluate_steps(host_env, 400) if mode == "steps" else rt.evaluate_episodes(host_env, 25)
    assert len(res_fused[1]) > 0
    assert [float(x) for x in res_fused[1]] == [float(x) for x in res_host[1]]
    assert float(res_fused[0]) == float(res_host[0])


def test_save_load_round_trip(tmp_path):
    from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase

    algo, _ = _trained(32, 10)
    path = os.path.join(tmp_path, "table.npy")
    algo.save(path)
    assert np.array_equal(np.load(path), np.asarray(algo.q_table))
    other = OptimalQLearningBase(19683, 9, 0.99, seed=0)
    other.load(path)
    states = np.arange(0, 19683, 997, dtype=np.int32)
    assert np.array_equal(other.get_states_q_values(states), algo.get_states_q_values(states))
    # a float64 table written by the reference's save() is accepted
    np.save(os.path.join(tmp_path, "ref64.npy"), np.asarray(algo.q_table, dtype=np.float64))
    other.load(os.path.join(tmp_path, "ref64"))
    assert np.array_equal(np.asarray(other.q_table), np.asarray(algo.q_table))


@pytest.mark.parametrize("runtime,extra", [("single_thread", []), ("parallel", ["--processes", "2"]), ("distributed", [])])
def test_throughput_benchmark_cli(tmp_path, runtime, extra):
    from dist_classicrl_b200.benchmarks import throughput_benchmark as tb

    tb.main(["--runtime", runtime, "--agents", "16", "--steps", "64", "--output-dir", str(tmp_path), *extra])
    files = [f for f in os.listdir(tmp_path) if f.startswith(runtime)]
    assert len(files) == 1
    res = json.load(open(os.path.join(tmp_path, files[0])))
    # the reference's schema (TPB:251-259, 310-317)
    for key in ("runtime", "total_steps", "effective_steps", "elapsed_time", "throughput", "step_multiplier", "timestamp",
                "num_agents", "num_processes", "mpi_rank", "mpi_size"):
        assert key in res
    assert res["effective_steps"] == 64 * 16 and res["throughput"] > 0 and "roofline" in res
