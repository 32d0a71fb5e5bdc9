from dist_classicrl_b200.algorithms.runtime.base_runtime import BaseRuntime
from dist_classicrl_b200.algorithms.runtime.parallel_runtime import ParallelQLearning
from dist_classicrl_b200.algorithms.runtime.q_learning_async_dist import DistAsyncQLearning
from dist_classicrl_b200.algorithms.runtime.single_thread_runtime import SingleThreadQLearning

__all__ = ["BaseRuntime", "DistAsyncQLearning", "ParallelQLearning", "SingleThreadQLearning"]
