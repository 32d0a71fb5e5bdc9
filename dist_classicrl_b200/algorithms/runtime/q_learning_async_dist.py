"""Actor/learner trainer (drop-in for ``algorithms/runtime/q_learning_async_dist.py``, MPI:28-447).

Same roles and message protocol as the reference's MPI trainer, on ``torch.distributed`` (one process per rank) or
on in-process queues (ranks as threads) instead of ``mpi4py``:

* rank 0 is the *master*: it owns the table (on its B200), answers every worker's observation with actions
  (``_choose_actions``, MPI:208-210), turns consecutive observations of a worker into transitions and queues them
  (MPI:216-268), keeps the episode bookkeeping (MPI:273-279), and runs the learner in a second thread
  (``update_q_table``, MPI:59-161) which applies the queued transitions in queue order in batches of at most
  ``batch_size`` and validates every ``val_every_n_steps`` transitions;
* ranks >= 1 are *workers*: they step their environment with the actions they are sent and report
  ``(next_states, rewards, terminateds, truncateds, infos)`` (``run_environment``, MPI:283-346).

What differs from the reference is mechanical: a worker's vector step travels through the experience queue as ONE
record of arrays (the reference enqueues one Python tuple per agent, MPI:237-268) and the learner slices the
concatenated records into batches, so the order in which transitions are applied -- agent order within a worker
step, worker steps in arrival order -- is the reference's, and the table work of both threads is the engine's
``choose_actions`` / ``learn`` kernels (the C ABI serialises the two threads on the table, SURVEY 8b "Threading").
"""

from __future__ import annotations

import logging
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Any

import numpy as np

from dist_classicrl_b200.algorithms.runtime.base_runtime import BaseRuntime

logger = logging.getLogger(__name__)

MASTER_RANK = 0
TAG_STOP, TAG_DATA = 0, 1  # MPI:281 (tag 0 = terminate), MPI:210 (tag 1 = actions)


# ------------------------------------------------------------------------------------------------ messengers
class Messenger:
    """Point-to-point messages between the ranks of one training job (the slice of ``MPI.COMM_WORLD`` the reference
    uses: ``send`` / ``recv`` / ``Iprobe`` / ``Barrier``)."""

    rank: int = 0
    size: int = 1

    def send(self, obj: Any, dest: int, tag: int = TAG_DATA) -> None:
        raise NotImplementedError

    def recv(self, source: int) -> tuple[Any, int]:
        """Next message from ``source`` and its tag (blocks)."""
        raise NotImplementedError

    def poll(self, source: int) -> bool:
        """True if ``recv(source)`` would not block (transports that cannot probe say True: the master then serves
        the workers round-robin, which is one of the arrival orders the reference admits)."""
        return True

    def barrier(self) -> None:
        raise NotImplementedError


class ThreadMessenger(Messenger):
    """Ranks as threads of one process (tests, single-node runs that keep every environment next to the GPU)."""

    def __init__(self, rank: int, size: int, boxes: dict, barrier: threading.Barrier) -> None:
        self.rank, self.size, self._boxes, self._barrier = rank, size, boxes, barrier

    @classmethod
    def group(cls, size: int) -> list["ThreadMessenger"]:
        boxes = {(s, d): queue.Queue() for s in range(size) for d in range(size) if s != d}
        bar = threading.Barrier(size)
        return [cls(r, size, boxes, bar) for r in range(size)]

    def send(self, obj, dest, tag=TAG_DATA):
        self._boxes[(self.rank, dest)].put((obj, tag))

    def recv(self, source):
        return self._boxes[(source, self.rank)].get()

    def poll(self, source):
        return not self._boxes[(source, self.rank)].empty()

    def barrier(self):
        self._barrier.wait()


class TorchDistMessenger(Messenger):
    """Ranks as ``torch.distributed`` processes (``gloo`` or ``nccl`` default group; objects are pickled like
    mpi4py's lower-case ``send`` / ``recv``)."""

    def __init__(self, group=None) -> None:
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("DistAsyncQLearning needs an initialised torch.distributed process group (or a messenger)")
        self._dist, self._group = dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)

    def send(self, obj, dest, tag=TAG_DATA):
        self._dist.send_object_list([(obj, tag)], dst=dest, group=self._group)

    def recv(self, source):
        box = [None]
        self._dist.recv_object_list(box, src=source, group=self._group)
        return box[0]

    def barrier(self):
        self._dist.barrier(group=self._group)


# ------------------------------------------------------------------------------------------------ trainer
def _obs_of(states):
    return states["observation"] if isinstance(states, dict) else states


class DistAsyncQLearning(BaseRuntime):
    """Distributed asynchronous Q-learning: workers run environments, the master selects actions and learns."""

    num_agents: int
    experience_queue: queue.Queue
    batch_size: int

    def __init__(self, algorithm, lr_schedule, exploration_rate_schedule, messenger: Messenger | None = None) -> None:
        super().__init__(algorithm, lr_schedule, exploration_rate_schedule)
        self._messenger = messenger

    @property
    def messenger(self) -> Messenger:
        if self._messenger is None:
            self._messenger = TorchDistMessenger()
        return self._messenger

    def init_training(self) -> None:
        return None

    def run_steps(self) -> None:  # noqa: D102  (the reference's signature differs from the ABC's too, MPI:53)
        return None

    def close_training(self) -> None:
        return None

    # -------------------------------------------------------------------------------------------- learner thread
    def update_q_table(self, val_env, val_every_n_steps: int, val_steps: int | None, val_episodes: int | None) -> list[float]:
        """Consume the experience queue until the ``None`` sentinel: learn in batches of at most ``batch_size``
        transitions in queue order, validate every ``val_every_n_steps`` transitions (MPI:59-161)."""
        running = True
        val_reward_history: list[float] = []
        pending: list[tuple] = []  # records not yet (fully) learned: (obs, actions, rewards, next_obs, next_masks, terminated)
        have = 0
        steps_since_val = 0
        step = 0
        while running or have:
            # one batch per turn: up to batch_size transitions, cut short at the validation boundary, by an empty queue
            # (MPI:103-108) or by the sentinel
            want = min(self.batch_size, val_every_n_steps - steps_since_val)
            while running and have < want:
                try:
                    rec = self.experience_queue.get(timeout=0.1)
                except queue.Empty:
                    break
                if rec is None:
                    running = False
                    break
                pending.append(rec)
                have += len(rec[0])
            take = min(have, want)
            if take:
                batch, pending = _take(pending, take)
                have -= take
                self._learn_batch(batch)
                steps_since_val += take
                step += take
            if steps_since_val >= val_every_n_steps:
                if val_steps is not None:
                    total, _agents = self.evaluate_steps(val_env, val_steps)
                else:
                    total, _agents = self.evaluate_episodes(val_env, val_episodes)
                val_reward_history.append(total)
                steps_since_val = 0
                logger.debug("Step %d, Eval total rewards: %s", step, total)
        return val_reward_history

    def _learn_batch(self, batch) -> None:
        obs, actions, rewards, next_obs, next_masks, terminated = batch
        if next_masks is None:
            self._learn(obs, actions, rewards, next_obs, terminated)
        else:  # MPI:123-139: states travel as {"observation": ...}, next states carry the masks
            self._learn({"observation": obs, "action_mask": None}, actions, rewards,
                        {"observation": next_obs, "action_mask": next_masks}, terminated)

    # -------------------------------------------------------------------------------------------- master
    def communicate_master(self, steps: int) -> list[float]:
        """Serve the workers until ``steps`` worker vector steps have been turned into transitions (MPI:163-281)."""
        msg = self.messenger
        workers = list(range(1, msg.size))
        reward_history: list[float] = []
        returns: dict[int, np.ndarray] = {}
        prev_states: dict[int, Any] = {w: None for w in workers}
        prev_actions: dict[int, Any] = {w: None for w in workers}
        step = 0
        while step < steps and workers:
            progressed = False
            for w in workers:
                if not msg.poll(w):
                    continue
                progressed = True
                data, _tag = msg.recv(w)
                assert data is not None, "Received None from worker"
                next_states, rewards, terminateds, truncateds, _infos = data
                actions = self._choose_actions(next_states)
                msg.send(actions, w, TAG_DATA)
                rewards = np.asarray(rewards, dtype=np.float32)
                if prev_states[w] is None:
                    returns[w] = np.zeros(len(rewards), dtype=np.float32)
                else:
                    step += 1
                    returns[w] += rewards
                    masks = next_states["action_mask"] if isinstance(next_states, dict) else None
                    self.experience_queue.put((
                        np.asarray(_obs_of(prev_states[w]), dtype=np.int32), np.asarray(prev_actions[w], dtype=np.int32), rewards,
                        np.asarray(_obs_of(next_states), dtype=np.int32), None if masks is None else np.asarray(masks, dtype=np.int32),
                        np.asarray(terminateds, dtype=bool)))
                prev_states[w], prev_actions[w] = next_states, actions
                done = np.logical_or(np.asarray(terminateds, dtype=bool), np.asarray(truncateds, dtype=bool))
                for idx in np.nonzero(done)[0]:
                    reward_history.append(returns[w][idx])
                    returns[w][idx] = 0
                if step >= steps:
                    break
            if not progressed:
                threading.Event().wait(0.0002)
        self.experience_queue.put(None)
        # every worker has exactly one message in flight (its reset observation, or the answer to the last actions it was
        # sent): drain those, then stop the workers (MPI:275-281)
        for w in workers:
            msg.recv(w)
        for w in workers:
            msg.send(None, w, TAG_STOP)
        return reward_history

    # -------------------------------------------------------------------------------------------- worker
    def run_environment(self, env, curr_state_dict: dict | None = None):
        """Step ``env`` with the master's actions until told to stop (MPI:283-346)."""
        msg = self.messenger
        msg.barrier()
        n = env.num_agents if hasattr(env, "num_agents") else env.num_envs
        if not curr_state_dict:
            states, infos = env.reset()
        else:
            states, infos = curr_state_dict["states"], curr_state_dict["infos"]
        rewards = np.zeros(n, dtype=np.float32)
        msg.send((states, rewards, np.zeros(n, dtype=bool), np.zeros(n, dtype=bool), infos), MASTER_RANK)
        while True:
            actions, tag = msg.recv(MASTER_RANK)
            if tag == TAG_STOP:
                break
            states, rewards, terminated, truncated, infos = env.step(actions)
            msg.send((states, rewards, terminated, truncated, infos), MASTER_RANK)
        msg.barrier()
        return env, {"states": states, "infos": infos, "rewards": rewards}

    # -------------------------------------------------------------------------------------------- entry point
    def train(self, env, steps: int, val_env, val_every_n_steps: int, val_steps: int | None, val_episodes: int | None,
              curr_state_dict: dict[str, Any] | None = None, *, batch_size: int = 32):
        """Master: ``(reward history, validation history, None, None)``; workers: ``([], [], env, state dict)``
        (MPI:348-447)."""
        assert (val_steps is None) ^ (val_episodes is None), "Either val_steps or val_episodes should be provided."
        msg = self.messenger
        if msg.rank == MASTER_RANK:
            msg.barrier()
            self.experience_queue = queue.Queue(maxsize=-1)
            self.batch_size = batch_size
            with ThreadPoolExecutor(max_workers=1) as executor:
                future = executor.submit(self._learner_entry, val_env, val_every_n_steps, val_steps, val_episodes)
                reward_history = self.communicate_master(steps)
                val_reward_history = future.result()
            msg.barrier()
            return reward_history, val_reward_history, None, None
        env, state = self.run_environment(env, curr_state_dict)
        return [], [], env, state

    def _learner_entry(self, val_env, val_every_n_steps, val_steps, val_episodes):
        device = getattr(self.algorithm, "device", None)
        if device is not None:
            import torch

            torch.cuda.set_device(device)  # the learner thread starts without a current device
        return self.update_q_table(val_env, val_every_n_steps, val_steps, val_episodes)


def _take(pending: list[tuple], count: int):
    """First ``count`` transitions of the queued records as one batch of arrays, and the records that remain."""
    parts: list[tuple] = []
    rest = list(pending)
    need = count
    while need:
        rec = rest[0]
        m = len(rec[0])
        if m <= need:
            parts.append(rec)
            rest.pop(0)
            need -= m
        else:
            parts.append(tuple(None if x is None else x[:need] for x in rec))
            rest[0] = tuple(None if x is None else x[need:] for x in rec)
            need = 0
    cols = []
    for c in range(6):
        xs = [p[c] for p in parts]
        cols.append(None if xs[0] is None else (xs[0] if len(xs) == 1 else np.concatenate(xs)))
    return tuple(cols), rest
