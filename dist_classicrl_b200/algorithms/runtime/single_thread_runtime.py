"""Single-process trainer (drop-in for ``algorithms/runtime/single_thread_runtime.py``, STR:22-79).

``run_steps`` keeps the reference's contract -- reset (or resume from ``curr_state_dict``), ``steps`` vector
steps, returns ``(mean episode reward, episode rewards, env, state dict)`` with the keys ``states, infos,
rewards, episode_rewards`` (STR:66-76) -- but with a GPU environment the whole loop is one fused kernel launch
per chunk instead of ``steps`` Python iterations.
"""

from __future__ import annotations

from typing import Any

import numpy as np

from dist_classicrl_b200.algorithms.runtime.base_runtime import BaseRuntime, _split


class SingleThreadQLearning(BaseRuntime):
    """Single-process Q-learning trainer."""

    def init_training(self) -> None:
        return None

    def close_training(self) -> None:
        return None

    def run_steps(self, steps: int, env, curr_state_dict: dict | None = None, *, trace: dict | None = None, _mean: bool = True):
        """``_mean=False`` (internal: callers that run a window of a longer ``run_steps``, e.g. ``ReplicatedQLearning``)
        skips the mean, so that a window in which no episode ends does not raise the reference's ZeroDivisionError
        (STR:67) in the middle of a run that does finish episodes."""
        reward_history: list[Any] = []
        if curr_state_dict is None:
            states, infos = env.reset()
            agent_rewards = np.zeros(len(_split(states)[0]), dtype=np.float32)  # STR:57
        else:
            states, infos, agent_rewards = curr_state_dict["states"], curr_state_dict["infos"], curr_state_dict["rewards"]
        if self._can_fuse(env):
            reward_history = self._run_fused(env, steps, agent_rewards, trace=trace)
            states = env._obs_lazy()
            if not _mean:
                mean = 0.0
            elif self.history_mode == "full":
                mean = sum(reward_history) / len(reward_history)  # ZeroDivisionError if no episode ended (STR:67)
            else:
                mean = self.last_episode_sum / self.last_episode_count if self.last_episode_count else 0.0
        else:
            for _ in range(steps):
                states, infos = self.run_single_step(env, states, agent_rewards, reward_history)
            mean = sum(reward_history) / len(reward_history) if _mean else 0.0
        return (
            mean,
            reward_history,
            env,
            {"states": states, "infos": infos, "rewards": agent_rewards, "episode_rewards": reward_history},
        )
