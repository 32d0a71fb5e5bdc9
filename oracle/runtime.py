"""Restatement of the single-thread training loop (TEST INFRASTRUCTURE).

``run_steps`` follows ``SingleThreadQLearning.run_steps`` (STR:28-76) driving
``BaseRuntime.run_single_step`` (BRT:184-222): select -> env.step ->
``agent_rewards += r`` -> learn -> schedules ``update(N)`` -> episode history in
agent order.  Schedules restate ``schedules/*.py`` (fp64 host scalars).
"""

from __future__ import annotations

import numpy as np

from oracle import qlearning as oq


class Constant:  # schedules/constant_schedule.py:6-12
    def __init__(self, value: float) -> None:
        self.value = value

    def get_value(self) -> float:
        return self.value

    def update(self, steps: int) -> None:
        pass


class Linear:  # schedules/linear_schedule.py:6-31
    def __init__(self, value: float, decay_rate: float) -> None:
        self.value, self.decay_rate = value, decay_rate

    def get_value(self) -> float:
        return self.value

    def update(self, steps: int) -> None:
        self.value = self.value + steps * self.decay_rate


class Exponential:  # schedules/exponential_schedule.py:6-31
    def __init__(self, value: float, min_value: float, decay_rate: float) -> None:
        self.value, self.min_value, self.decay_rate = value, min_value, decay_rate

    def get_value(self) -> float:
        return self.value

    def update(self, steps: int) -> None:
        self.value = max(self.value * (self.decay_rate**steps), self.min_value)


def run_steps(q, gamma, env, uniforms, lr_sched, eps_sched, *, t0=0, states=None, agent_rewards=None, record=False):
    """Run ``uniforms.shape[0]`` vector steps; returns ``(reward_history, states, agent_rewards, trace)``.

    ``uniforms[t]`` feeds step ``t0 + t``.  ``trace`` (if ``record``) holds per-step
    ``actions, rewards, terminated, next observation`` for bit-exact comparison.
    """
    if states is None:
        states, _ = env.reset()
    n = env.num_envs
    if agent_rewards is None:
        agent_rewards = np.zeros(n, dtype=np.float32)  # STR:57
    history: list[float] = []
    trace = {"actions": [], "rewards": [], "terminated": [], "obs": []} if record else None
    for t in range(uniforms.shape[0]):
        u = uniforms[t]
        is_dict = isinstance(states, dict)
        obs = states["observation"] if is_dict else states
        masks = states["action_mask"] if is_dict else None
        actions = oq.select(q, obs, masks, eps_sched.get_value(), u)  # BRT:208, 265-291
        nxt, rewards, term, trunc, _ = env.step(actions, u)  # BRT:210
        agent_rewards += rewards  # BRT:212
        nobs = nxt["observation"] if is_dict else nxt
        nmask = nxt["action_mask"] if is_dict else None
        oq.learn_sequential(q, obs, actions, rewards, nobs, term, lr_sched.get_value(), gamma, nmask)  # BRT:214
        lr_sched.update(n)  # BRT:262-263
        eps_sched.update(n)
        states = nxt
        for i in range(n):  # BRT:218-221
            if term[i] or trunc[i]:
                history.append(float(agent_rewards[i]))
                agent_rewards[i] = 0
        if record:
            trace["actions"].append(actions.copy())
            trace["rewards"].append(rewards.copy())
            trace["terminated"].append(term.copy())
            trace["obs"].append(np.asarray(nobs).copy())
    if record:
        trace = {k: np.stack(v) for k, v in trace.items()}
    return history, states, agent_rewards, trace
