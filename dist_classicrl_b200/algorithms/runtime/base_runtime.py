"""Trainer base class (drop-in for ``algorithms/runtime/base_runtime.py``, BRT).

Same constructor and public methods as the reference ``BaseRuntime`` (BRT:41-49, 55-93, 99-182, 184-222,
293-384).  Two ways to advance one vector step:

* the *unfused* path -- ``run_single_step`` = ``choose_actions`` -> ``env.step`` -> ``learn`` -> schedules
  ``update(N)`` -> episode bookkeeping, exactly the reference's sequence (BRT:208-221), usable with any host
  environment;
* the *fused* path -- when the environment is one of the engine's GPU environments
  (:class:`~dist_classicrl_b200.environments.custom_env.DeviceVecEnv`) and both random streams are engine
  streams, ``_run_fused`` executes K vector steps in one persistent cooperative kernel
  (``qe_fused_steps``) and returns the same history the loop would have produced.
"""

from __future__ import annotations

import ctypes as C
import logging
from abc import ABC, abstractmethod
from typing import Any

import numpy as np

from dist_classicrl_b200 import capi, hostmem
from dist_classicrl_b200.environments.custom_env import DeviceVecEnv
from dist_classicrl_b200.rng import PredrawnUniforms, explore_threshold, is_engine_rng

logger = logging.getLogger(__name__)

_TRACE_BYTES = 256 << 20  # per-chunk budget of the episode-return trace
_PIN_BYTES = 64 << 10  # host arrays at least this large are page-locked in place before they are copied


def _split(states):
    """(observation array, action-mask array or None) of a reset/step observation (BRT:239, 283)."""
    if isinstance(states, dict):
        return states["observation"], states["action_mask"]
    return states, None


def _book_episodes(agent_rewards, terminateds, truncateds, reward_history) -> None:
    """Agent-order episode bookkeeping of BRT:218-221."""
    done = np.logical_or(np.asarray(terminateds), np.asarray(truncateds))
    for i in np.nonzero(done)[0]:
        reward_history.append(agent_rewards[i])
        agent_rewards[i] = 0


class BaseRuntime(ABC):
    """Q-learning trainer: an algorithm plus a learning-rate and an exploration-rate schedule."""

    def __init__(self, algorithm, lr_schedule, exploration_rate_schedule) -> None:
        self.algorithm = algorithm
        self.lr_schedule = lr_schedule
        self.exploration_rate_schedule = exploration_rate_schedule
        self.history_mode = "full"  # "summary": keep only count/sum of finished episodes (huge agent counts)
        self.last_episode_count = 0
        self.last_episode_sum = 0.0
        # "sequential": the reference trainers' update, learn -> learn_iter (BRT:245-259, QLO:893-934).  "accumulate":
        # learn_vec instead (QLO:819-891; snapshot bootstrap + accumulating scatter = plain atomics on the GPU)
        self.td_update = "sequential"

    @abstractmethod
    def init_training(self) -> None:
        """Prepare for training."""

    @abstractmethod
    def run_steps(self, steps: int, env, curr_state_dict):
        """Run ``steps`` vector steps; returns ``(mean episode reward, episode rewards, env, state dict)``."""

    @abstractmethod
    def close_training(self) -> None:
        """Release training resources."""

    # ------------------------------------------------------------------ train (BRT:99-182)
    def train(self, env, steps: int, val_env, val_every_n_steps: int, val_steps: int | None = None,
              val_episodes: int | None = None, curr_state_dict: dict | None = None):
        assert (val_steps is None) ^ (val_episodes is None), "Exactly one of val_steps or val_episodes must be specified."
        self.init_training()
        reward_history: list[float] = []
        val_reward_history: list[float] = []
        state_dict = None
        done = 0
        while done < steps:
            chunk = min(val_every_n_steps, steps - done)
            # like the reference, every chunk starts from `curr_state_dict`, not from the previous chunk (BRT:156-161)
            _, episode_rewards, env, state_dict = self.run_steps(steps=chunk, env=env, curr_state_dict=curr_state_dict)
            reward_history.extend(episode_rewards)
            if val_steps is not None:
                total, _per_agent = self.evaluate_steps(val_env, val_steps)
            else:
                total, _per_agent = self.evaluate_episodes(val_env, val_episodes)
            val_reward_history.append(total)
            logger.debug("Step %d, Eval total rewards: %s", done + 1, total)
            done += val_every_n_steps
        self.close_training()
        return reward_history, val_reward_history, env, state_dict

    # ------------------------------------------------------------------ unfused vector step (BRT:184-291)
    def run_single_step(self, env, states, agent_rewards, reward_history):
        actions = self._choose_actions(states)
        next_states, rewards, terminateds, truncateds, infos = env.step(actions)
        agent_rewards += rewards
        self._learn(states, actions, rewards, next_states, terminateds)
        _book_episodes(agent_rewards, terminateds, truncateds, reward_history)
        return next_states, infos

    def _learn(self, states, actions, rewards, next_states, terminateds) -> None:
        obs, _ = _split(states)
        next_obs, next_masks = _split(next_states)
        assert isinstance(states, dict) == isinstance(next_states, dict)
        lr = self.lr_schedule.get_value()
        learn = self.algorithm.learn_vec if self.td_update == "accumulate" else self.algorithm.learn
        if next_masks is None:
            learn(obs, actions, rewards, next_obs, terminateds, lr)
        else:
            learn(obs, actions, rewards, next_obs, terminateds, lr, next_masks)
        n_updates = len(obs)
        self.lr_schedule.update(n_updates)
        self.exploration_rate_schedule.update(n_updates)

    def _choose_actions(self, states):
        obs, masks = _split(states)
        eps = self.exploration_rate_schedule.get_value()
        if masks is None:
            return self.algorithm.choose_actions(obs, exploration_rate=eps)
        return self.algorithm.choose_actions(states=obs, action_masks=masks, exploration_rate=eps)

    # ------------------------------------------------------------------ evaluation (BRT:293-384)
    def _greedy_actions(self, states):
        obs, masks = _split(states)
        if masks is None:
            return self.algorithm.choose_actions(obs, exploration_rate=0.0, deterministic=True)
        return self.algorithm.choose_actions(states=obs, action_masks=masks, exploration_rate=0.0, deterministic=True)

    def _evaluate_fused(self, env, steps: int | None, episodes: int | None):
        """Evaluation on the device: greedy select + env step in the fused loop, table untouched."""
        n = env.num_envs
        agent_rewards = np.zeros(n, dtype=np.float32)
        history: list[float] = []
        saved = self.history_mode
        self.history_mode = "full"
        try:
            if steps is not None:
                vector_steps = len(range(0, steps, n))
                if vector_steps:
                    history = self._run_fused(env, vector_steps, agent_rewards, evaluate=True)
            else:
                while len(history) < episodes:  # whole vector steps, like BRT:375-383: stop after the step that reaches the count
                    trace: dict = {}
                    self._run_fused(env, 16, agent_rewards, evaluate=True, trace=trace, history_rows=True)
                    for row in trace["episode_rows"]:
                        history.extend(row)
                        if len(history) >= episodes:
                            break
        finally:
            self.history_mode = saved
        return sum(history), history

    def evaluate_steps(self, env, steps: int):
        if self._can_fuse(env) and not isinstance(self.algorithm._rng, PredrawnUniforms):
            env.reset(seed=42)
            return self._evaluate_fused(env, steps, None)
        states, _ = env.reset(seed=42)
        n_agents = len(_split(states)[0])
        agent_rewards = np.zeros(n_agents, dtype=np.float32)
        reward_history: list[float] = []
        for _ in range(0, steps, n_agents):
            states, rewards, terminateds, truncateds, _infos = env.step(self._greedy_actions(states))
            agent_rewards += rewards
            _book_episodes(agent_rewards, terminateds, truncateds, reward_history)
        return sum(reward_history), reward_history

    def evaluate_episodes(self, env, episodes: int):
        if self._can_fuse(env) and not isinstance(self.algorithm._rng, PredrawnUniforms):
            env.reset(seed=42)
            return self._evaluate_fused(env, None, episodes)
        states, _ = env.reset(seed=42)
        n_agents = len(_split(states)[0])
        agent_rewards = np.zeros(n_agents, dtype=np.float32)
        reward_history: list[float] = []
        while len(reward_history) < episodes:
            states, rewards, terminateds, truncateds, _infos = env.step(self._greedy_actions(states))
            agent_rewards += rewards
            _book_episodes(agent_rewards, terminateds, truncateds, reward_history)
        return sum(reward_history), reward_history

    # ------------------------------------------------------------------ fused path
    def _can_fuse(self, env) -> bool:
        algo = self.algorithm
        if not isinstance(env, DeviceVecEnv) or getattr(env, "output", "numpy") == "torch-unfused":
            return False
        if not (is_engine_rng(algo._rng) and is_engine_rng(env._rng)):
            return False
        if isinstance(algo._rng, PredrawnUniforms) != isinstance(env._rng, PredrawnUniforms):
            return False
        if isinstance(algo._rng, PredrawnUniforms) and (algo._rng.uniforms is not env._rng.uniforms or algo._rng.t != env._rng.t):
            return False
        return algo.action_size <= 32 and env.num_actions == algo.action_size and env.num_states == algo.state_size

    def _uniforms_on_device(self, rng, block, t0: int, k: int, n: int, dev):
        """Device copy of the pre-drawn uniforms ``[t0, t0+k)``.  The block after it is copied on a side stream right
        away (asynchronously when the host array is pinned), so the next call's host-to-device copy overlaps this
        call's kernel."""
        import torch

        main = torch.cuda.current_stream()
        cached = getattr(rng, "_prefetched", None)
        if cached is not None and cached[0] == (t0, k, n, str(dev)):
            main.wait_event(cached[2])
            u_dev = cached[1]
        else:
            u_dev = torch.from_numpy(block.view(np.int32)).to(dev)
        rng._prefetched = None
        nxt = rng.uniforms[t0 + k:t0 + 2 * k, :n]
        if nxt.shape[0] == k and nxt.flags.c_contiguous:
            src = torch.from_numpy(nxt.view(np.int32))
            if src.is_pinned():
                side = getattr(self, "_copy_stream", None)
                if side is None:
                    side = self._copy_stream = torch.cuda.Stream(device=dev)
                with torch.cuda.stream(side):
                    ahead = torch.empty(src.shape, dtype=torch.int32, device=dev)
                    ahead.copy_(src, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(side)
                ahead.record_stream(main)
                rng._prefetched = ((t0 + k, k, n, str(dev)), ahead, done)
        return u_dev

    def _run_fused(self, env, steps: int, agent_rewards, *, trace: dict | None = None, evaluate: bool = False,
                   history_rows: bool = False) -> list[float]:
        """K vector steps in ``qe_fused_steps`` launches; returns the episode-reward history (agent order
        within a step, steps in order -- BRT:218-221)."""
        import torch

        algo = self.algorithm
        lib = capi.lib()
        algo._before_device_op()
        n = env.num_envs
        dev = env.device
        # the caller's running returns (a plain NumPy array in the reference's state dictionary, STR:57) are page-locked
        # in place on first sight: the per-call upload and the copy back are then asynchronous DMA on this stream
        host_ret = None
        if isinstance(agent_rewards, torch.Tensor):
            ep_ret = agent_rewards
        else:
            if (isinstance(agent_rewards, np.ndarray) and agent_rewards.dtype == np.float32 and agent_rewards.shape == (n,)
                    and agent_rewards.nbytes >= _PIN_BYTES and hostmem.pin(agent_rewards)):
                host_ret = torch.from_numpy(agent_rewards)
                ep_ret = torch.empty(n, dtype=torch.float32, device=dev)
                ep_ret.copy_(host_ret, non_blocking=True)
            else:
                ep_ret = torch.from_numpy(np.ascontiguousarray(agent_rewards, dtype=np.float32)).to(dev)
        variant = algo._variant(n, False, env.dict_obs)
        full = self.history_mode == "full"
        chunk = max(1, min(steps, _TRACE_BYTES // (4 * n))) if full else steps
        ep_stats = torch.zeros(2, dtype=torch.int64, device=dev)  # [0] = float64 sum of finished episodes, [1] = their count
        history: list[float] = []
        predrawn = isinstance(algo._rng, PredrawnUniforms)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        ag = env.agents_struct(ep_ret)
        done = 0
        stats_host = torch.zeros(2, dtype=torch.int64)
        while done < steps:
            k = min(chunk, steps - done)
            th = np.empty(k, dtype=np.uint64)
            lrs = np.empty(k, dtype=np.float32)
            for j in range(k):  # BRT:245-263: values read before the update of the same step
                if evaluate:  # deterministic=True, exploration_rate=0.0 (BRT:319-328); schedules untouched
                    th[j], lrs[j] = 0, 0.0
                    continue
                th[j] = explore_threshold(self.exploration_rate_schedule.get_value())
                lrs[j] = np.float32(self.lr_schedule.get_value())
                self.lr_schedule.update(n)
                self.exploration_rate_schedule.update(n)
            run = capi.QeRun()
            run.steps = k
            run.evaluate = int(evaluate)
            run.learn_mode = capi.QE_LEARN_ACCUMULATE if self.td_update == "accumulate" else capi.QE_LEARN_SEQUENTIAL
            run.explore_thresholds_host = th.ctypes.data_as(C.c_void_p)
            run.learning_rates_host = lrs.ctypes.data_as(C.c_void_p)
            u_dev = None
            if predrawn:
                t0 = algo._rng.t
                block = np.ascontiguousarray(algo._rng.uniforms[t0:t0 + k, :n])
                if block.shape[0] < k:
                    raise IndexError("pre-drawn uniform stream exhausted")
                u_dev = self._uniforms_on_device(algo._rng, block, t0, k, n, dev)
                run.uniforms = u_dev.data_ptr()
                run.slots = block.shape[2]
                algo._rng.t += k
                if env._rng is not algo._rng:
                    env._rng.t += k
            else:
                run.slots = env.slots
                run.stream_seed, run.t0 = algo._rng.seed, algo._rng.t
                run.env_stream_seed, run.env_t0 = env._rng.seed, env._rng.t
                algo._rng.t = (algo._rng.t + k) & 0xFFFFFFFF
                if env._rng is not algo._rng:
                    env._rng.t = (env._rng.t + k) & 0xFFFFFFFF
            run.agent0 = getattr(env, "agent0", 0)
            run.empty_all = int(variant != "iter")
            run.use_masks = int(env.dict_obs)
            tr_ep = None
            if full:
                tr_ep = torch.empty((k, n), dtype=torch.float32, device=dev)
                run.trace_episode_returns = tr_ep.data_ptr()
            keep = []
            if trace is not None:
                for name, field, dt in (("actions", "trace_actions", torch.int32), ("rewards", "trace_rewards", torch.float32),
                                        ("terminated", "trace_terminated", torch.uint8), ("obs", "trace_next_states", torch.int32)):
                    buf = torch.empty((k, n), dtype=dt, device=dev)
                    setattr(run, field, buf.data_ptr())
                    keep.append((name, buf))
            run.episode_sum, run.episode_count = ep_stats.data_ptr(), ep_stats.data_ptr() + 8
            capi.check(lib.qe_fused_steps(algo.handle, C.byref(ag), C.byref(run), stream))
            done += k
            if done == steps:  # the results travel back behind the last launch, before the one synchronisation
                stats_host = self._stats_host()
                stats_host.copy_(ep_stats, non_blocking=True)
                if host_ret is not None:
                    host_ret.copy_(ep_ret, non_blocking=True)
            capi.check(lib.qe_sync(algo.handle, stream))
            if full:
                flat = tr_ep.reshape(-1)
                history.extend(flat[~torch.isnan(flat)].cpu().tolist())
                if history_rows and trace is not None:
                    host = tr_ep.cpu().numpy()
                    trace.setdefault("episode_rows", []).extend(row[~np.isnan(row)].tolist() for row in host)
            for name, buf in keep:
                trace.setdefault(name, []).append(buf.cpu().numpy())
        if not evaluate:
            algo._device_wrote()
        env.refresh_after_fused()
        self.last_episode_count = int(stats_host[1])
        self.last_episode_sum = float(stats_host[:1].view(torch.float64)[0])
        if host_ret is None and isinstance(agent_rewards, np.ndarray):
            agent_rewards[:] = ep_ret.cpu().numpy()
        return history

    def _stats_host(self):
        """Page-locked landing buffer of the fused loop's episode statistics."""
        import torch

        buf = getattr(self, "_stats_pinned", None)
        if buf is None:
            buf = self._stats_pinned = torch.zeros(2, dtype=torch.int64).pin_memory()
        return buf
