"""GPU counterpart of the reference's ``benchmarks/throughput_benchmark.py`` (TPB:105-323).

Same command line (``--runtime {single_thread,parallel,distributed} --agents N --processes P --steps T
--output-dir DIR [--verbose]``), same workload (TicTacToe against the random machine, gamma 0.99, lr
Exp(0.1, 1e-5, 0.995), eps Exp(1.0, 0.01, 0.995), validation disabled, TPB:53-58, 205-222) and the same JSON file
names and keys (TPB:251-259, 310-317), so the reference's ``analyze_throughput_benchmarks.py`` can read GPU results
next to its published CPU ones.  Extra keys: ``device``, ``n_gpus`` and a ``roofline`` block.

``distributed`` = one rank per GPU under ``python -m torch.distributed.run`` (replicated table, Q-delta all-reduce);
with a single process it degenerates to one GPU, like the reference run without ``mpirun``.
"""

from __future__ import annotations

import argparse
import json
import logging
import os
import time
from pathlib import Path
from typing import Any

logger = logging.getLogger(__name__)

DEFAULT_STEPS = 100000
LEARNING_RATE, DISCOUNT, EXPLORATION, DECAY, MIN_EXPLORATION = 0.1, 0.99, 1.0, 0.995, 0.01
RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))


def is_master(runtime: str) -> bool:
    return runtime != "distributed" or RANK == 0


def create_environments(num_agents: int, num_processes: int, runtime: str):
    from dist_classicrl_b200.environments import TicTacToeVecEnv

    val_env = TicTacToeVecEnv(1, seed=42)
    if runtime == "parallel":
        return [TicTacToeVecEnv(num_agents, seed=k) for k in range(num_processes)], val_env
    env = TicTacToeVecEnv(num_agents, seed=RANK)
    env.agent0 = RANK * num_agents
    return env, val_env


def initialize_agent(env, runtime: str):
    from dist_classicrl_b200 import spaces
    from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase
    from dist_classicrl_b200.algorithms.runtime import ParallelQLearning, SingleThreadQLearning
    from dist_classicrl_b200.schedules import ExponentialSchedule

    ref_env = env[0] if isinstance(env, (list, tuple)) else env
    assert isinstance(ref_env.single_action_space, spaces.Discrete)
    obs_space = ref_env.single_observation_space
    state_size = obs_space.spaces["observation"].n if isinstance(obs_space, spaces.Dict) else obs_space.n
    algo = OptimalQLearningBase(state_size=state_size, action_size=ref_env.single_action_space.n, discount_factor=DISCOUNT)
    lr = ExponentialSchedule(value=LEARNING_RATE, min_value=1e-5, decay_rate=DECAY)
    eps = ExponentialSchedule(value=EXPLORATION, min_value=MIN_EXPLORATION, decay_rate=DECAY)
    if runtime == "parallel":
        return ParallelQLearning(algo, lr, eps)
    if runtime in ("single_thread", "distributed"):
        return SingleThreadQLearning(algo, lr, eps)
    raise ValueError(f"Unknown runtime type: {runtime}")


def run_benchmark(agent, env, val_env, total_steps: int, runtime: str) -> dict[str, Any]:
    import torch

    tp = None
    if runtime == "distributed" and WORLD > 1:
        import torch.distributed as dist

        from dist_classicrl_b200.distributed import ReplicatedQLearning, TorchDistTransport

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if not dist.is_initialized():
            dist.init_process_group("nccl")
        tp = TorchDistTransport()
    agent.history_mode = "summary"
    torch.cuda.synchronize()
    start = time.perf_counter()
    if tp is not None:
        ReplicatedQLearning(agent, tp, sync_every=64).run_steps(total_steps, env)
    else:
        agent.train(env=env, steps=total_steps, val_env=val_env, val_every_n_steps=total_steps * 2, val_steps=None,
                    val_episodes=10, curr_state_dict=None)
    torch.cuda.synchronize()
    elapsed = time.perf_counter() - start
    ref_env = env[0] if isinstance(env, (list, tuple)) else env
    step_multiplier = ref_env.num_envs
    effective = total_steps * step_multiplier * (WORLD if tp is not None else 1)
    algo = agent.algorithm
    balg = 8 * algo.action_size + 12  # two row reads, one cell write, agent state read + write
    return {
        "runtime": runtime, "total_steps": total_steps, "effective_steps": effective, "elapsed_time": elapsed,
        "throughput": effective / elapsed, "step_multiplier": step_multiplier, "timestamp": time.time(),
        "device": torch.cuda.get_device_name(), "n_gpus": WORLD if tp is not None else 1,
        "roofline": {"bound": "launch/latency at this size (SURVEY 7.3-5); HBM for large tables", "algorithmic_bytes_per_agent_step": balg,
                     "achieved_gbs": effective * balg / elapsed / 1e9},
    }


def save_results(results: dict[str, Any], num_agents: int, num_processes: int, output_dir: str = "benchmark_results") -> str:
    out = Path(output_dir)
    out.mkdir(exist_ok=True)
    runtime = results["runtime"]
    if runtime == "single_thread":
        name = f"{runtime}_{num_agents}_agents.json"
    elif runtime == "parallel":
        name = f"{runtime}_{num_agents}_agents_{num_processes}_processes.json"
    else:
        name = f"{runtime}_{num_agents}_agents_{WORLD}_processes.json"
    results.update({"num_agents": num_agents, "num_processes": num_processes, "mpi_rank": RANK, "mpi_size": WORLD})
    path = out / name
    with path.open("w") as f:
        json.dump(results, f, indent=2)
    return str(path)


def parse_arguments(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Benchmark throughput of the B200 Q-learning runtimes.")
    p.add_argument("--runtime", choices=["single_thread", "parallel", "distributed"], required=True)
    p.add_argument("--agents", type=int, default=10, help="agents per vectorized environment (default: 10)")
    p.add_argument("--processes", type=int, default=1, help="environments sharing the table in the parallel runtime (default: 1)")
    p.add_argument("--steps", type=int, default=DEFAULT_STEPS, help=f"training steps (default: {DEFAULT_STEPS})")
    p.add_argument("--output-dir", default="benchmark_results")
    p.add_argument("--verbose", action="store_true")
    return p.parse_args(argv)


def main(argv=None) -> None:
    args = parse_arguments(argv)
    logging.basicConfig(level=logging.DEBUG if args.verbose else logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")
    env, val_env = create_environments(args.agents, args.processes, args.runtime)
    agent = initialize_agent(env, args.runtime)
    results = run_benchmark(agent, env, val_env, args.steps, args.runtime)
    if is_master(args.runtime):
        path = save_results(results, args.agents, args.processes, args.output_dir)
        logger.info("Throughput: %.2f steps/second; results saved to %s", results["throughput"], path)


if __name__ == "__main__":
    main()
