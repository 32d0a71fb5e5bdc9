"""Development aid: where the host time of the e2e arm (run_steps(1) per vector step, pre-drawn uniforms) goes."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase  # noqa: E402
from dist_classicrl_b200.algorithms.runtime import SingleThreadQLearning  # noqa: E402
from dist_classicrl_b200.environments import HashMDPVecEnv  # noqa: E402
from dist_classicrl_b200.rng import PredrawnUniforms, draw_uniforms  # noqa: E402
from dist_classicrl_b200.schedules import ConstantSchedule  # noqa: E402

s, a, n, steps = 1_000_000, 16, 1 << 20, 60
algo = OptimalQLearningBase(s, a, 0.99, seed=0, device=0)
algo.fill_random(1)
env = HashMDPVecEnv(n, s, a, env_seed=0, p_term=0.05, seed=0, device=0, output="torch")
env.attach(algo)
env.reset()
rt = SingleThreadQLearning(algo, ConstantSchedule(0.1), ConstantSchedule(0.1))
rt.history_mode = "summary"
u_host = torch.empty((steps, n, 4), dtype=torch.int32).pin_memory()
u_host.numpy().view(np.uint32)[:] = draw_uniforms(0, 0, steps, n, 4)
algo._rng = env._rng = PredrawnUniforms(u_host.numpy().view(np.uint32))
sd = {"states": None, "infos": {}, "rewards": np.zeros(n, dtype=np.float32)}
for _ in range(20):
    _, _, _, sd = rt.run_steps(1, env, sd)
torch.cuda.synchronize()
t = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    _, _, _, sd = rt.run_steps(1, env, sd)
pr.disable()
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t) / 30 * 1e3:.3f} ms per step (with profiler)")
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
