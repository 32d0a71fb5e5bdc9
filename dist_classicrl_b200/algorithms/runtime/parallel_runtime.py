"""Multi-environment trainer with one shared table (drop-in for ``algorithms/runtime/parallel_runtime.py``, PRT).

The reference starts one OS process per environment; all of them attach the Q-table through
``multiprocessing.shared_memory`` and take ONE global lock around ``choose_actions`` and around ``learn``
(PRT:250-258), so the algorithm work of the processes is serialised in some interleaving and only ``env.step``
overlaps.  On the B200 the table already lives in one place (HBM) and every environment's agents are stepped
by the same GPU, so the environments are advanced round-robin in blocks of ``interleave`` vector steps -- one
of the interleavings the reference's lock admits -- each block being one fused kernel launch.  As in the
reference every environment runs ``int(steps / len(env))`` vector steps (PRT:129), schedules are quantised to
fp32 (``set_mp``, PRT:65-66) and the per-environment episode histories are merged step-wise (PRT:157-158).
"""

from __future__ import annotations

from itertools import zip_longest
from typing import Any

import numpy as np

from dist_classicrl_b200.algorithms.runtime.base_runtime import BaseRuntime, _split


class ParallelQLearning(BaseRuntime):
    """Several environments, one device-resident table."""

    def __init__(self, *args: Any, interleave: int = 8, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.interleave = max(1, int(interleave))
        self.lr_schedule.set_mp()
        self.exploration_rate_schedule.set_mp()

    def init_training(self) -> None:
        self.algorithm._before_device_op()  # the table is "shared" as soon as it is on the device

    def close_training(self) -> None:
        return None

    def run_steps(self, steps: int, env, curr_state_dict: list[dict] | None):
        envs = list(env)
        per_env = int(steps / len(envs))
        slots = []
        for k, e in enumerate(envs):
            if curr_state_dict is None or curr_state_dict[k] is None:
                states, infos = e.reset()
                rewards = np.zeros(len(_split(states)[0]), dtype=np.float32)
            else:
                d = curr_state_dict[k]
                states, infos, rewards = d["states"], d["infos"], d["rewards"]
            slots.append({"states": states, "infos": infos, "rewards": rewards, "episode_rewards": []})
        done = 0
        while done < per_env:
            block = min(self.interleave, per_env - done)
            for e, slot in zip(envs, slots):
                if self._can_fuse(e):
                    slot["episode_rewards"].extend(self._run_fused(e, block, slot["rewards"]))
                    slot["states"] = e._obs_lazy()
                else:
                    for _ in range(block):
                        slot["states"], slot["infos"] = self.run_single_step(e, slot["states"], slot["rewards"],
                                                                            slot["episode_rewards"])
            done += block
        reward_history: list[float] = []
        for row in zip_longest(*[s["episode_rewards"] for s in slots]):
            reward_history.extend(r for r in row if r is not None)
        state_dicts = [{k: v for k, v in s.items() if k != "episode_rewards"} for s in slots]
        mean = sum(reward_history) / len(reward_history) if reward_history else 0.0
        return mean, reward_history, envs, state_dicts
