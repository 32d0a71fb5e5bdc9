"""Host logic of the actor/learner trainer (``DistAsyncQLearning``, reference MPI:28-447) on CPU: ranks as threads and
as gloo processes (world size 2 and 3).  The table on the master is a NumPy stand-in driven by the oracle's
``learn_sequential`` (the GPU tests run the same cases on the engine); the cases are the reference's own distributed
tests (``tests/dist_tests/test_q_learning_distributed.py:100-205``: ``Q[0,1] == 5.0``, eps schedule 6.0, who returns what)."""
import os
import socket
import threading

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from dist_classicrl_b200.algorithms.runtime import q_learning_async_dist as AD  # noqa: E402
from dist_classicrl_b200.schedules import ConstantSchedule, LinearSchedule  # noqa: E402
from oracle import qlearning as oq  # noqa: E402
from oracle.envs import BanditVec  # noqa: E402


class HostTable:
    """Duck-typed algorithm: always explores and picks action 1 (the reference tests' DeterministicRNG), learns with the
    oracle's sequential update; records the batches it was given."""

    def __init__(self, states=1, actions=2, gamma=1.0):
        self.q_table = np.zeros((states, actions), dtype=np.float32)
        self.discount_factor = gamma
        self.batches = []
        self.lock = threading.Lock()

    def choose_actions(self, states, exploration_rate=0.0, deterministic=False, action_masks=None):
        n = len(states)
        if deterministic:
            return np.argmax(self.q_table[np.asarray(states)], axis=1).astype(np.int32)
        return np.ones(n, dtype=np.int32)

    def learn(self, states, actions, rewards, next_states, terminated, lr, next_masks=None):
        with self.lock:
            self.batches.append(len(states))
            oq.learn_sequential(self.q_table, states, actions, rewards, next_states, terminated, lr, self.discount_factor, next_masks)


def _run_rank(msg, steps, episode_len, val_every, val_steps, batch, n_envs, out):
    algo = HostTable()
    rt = AD.DistAsyncQLearning(algo, ConstantSchedule(1.0), LinearSchedule(1.0, 1.0), messenger=msg)
    res = rt.train(env=BanditVec(n_envs, episode_len), steps=steps, val_env=BanditVec(1, episode_len), val_every_n_steps=val_every,
                   val_steps=val_steps, val_episodes=None, curr_state_dict={}, batch_size=batch)
    out[msg.rank] = (res, algo.q_table.copy(), rt.exploration_rate_schedule.get_value(), rt.lr_schedule.get_value(), list(algo.batches))


def _check_skip_validation(out, world):
    (hist, val, envs, state), q, eps, lr, batches = out[0]
    assert val == [] and isinstance(hist, list) and envs is None and state is None
    assert q.shape == (1, 2) and q[0, 1] == 5.0 and q[0, 0] == 0.0  # T-MPI:133-136
    assert lr == 1.0 and eps == 6.0                                  # T-MPI:138-139
    assert sum(batches) == 5 and max(batches) <= 8
    for r in range(1, world):
        (hist, val, env, state), *_ = out[r]
        assert hist == [] and val == [] and isinstance(env, BanditVec) and "states" in state


@pytest.mark.parametrize("world", [2, 3])
def test_thread_ranks_train_skips_validation_and_updates_q_table(world):  # T-MPI:100-147
    out = {}
    ms = AD.ThreadMessenger.group(world)
    ts = [threading.Thread(target=_run_rank, args=(m, 5, 6, 6, 6, 8, 1, out)) for m in ms]
    [t.start() for t in ts]
    [t.join(timeout=60) for t in ts]
    assert len(out) == world
    _check_skip_validation(out, world)


def test_thread_ranks_train_with_validation_collects_history():  # T-MPI:150-205
    out = {}
    ms = AD.ThreadMessenger.group(2)
    ts = [threading.Thread(target=_run_rank, args=(m, 6, 5, 3, 5, 2, 1, out)) for m in ms]
    [t.start() for t in ts]
    [t.join(timeout=60) for t in ts]
    (hist, val, envs, state), q, eps, _lr, batches = out[0]
    assert len(val) == 2                  # one validation per 3 transitions
    assert all(v == 5.0 for v in val)     # greedy policy pulls arm 1 for 5 steps
    assert hist == [5.0]                  # one finished episode of 5 steps
    assert sum(batches) == 6 and max(batches) <= 2
    assert eps == 7.0


def test_batches_follow_queue_order_and_size():
    """Several agents per worker: a worker's vector step is sliced into batches of at most batch_size in agent order."""
    out = {}
    ms = AD.ThreadMessenger.group(3)
    ts = [threading.Thread(target=_run_rank, args=(m, 8, 4, 1000, 1000, 3, 4, out)) for m in ms]
    [t.start() for t in ts]
    [t.join(timeout=60) for t in ts]
    (hist, _val, _e, _s), q, eps, _lr, batches = out[0]
    assert sum(batches) == 8 * 4 and max(batches) <= 3
    assert eps == 1.0 + 32.0
    assert len(hist) >= 4 and set(hist) == {4.0}


def test_nothing_queued_is_dropped_when_the_sentinel_arrives_early():
    """Validation every 5 transitions while 32 are queued in 8 records of 4: every transition is learned, 6 validations."""
    out = {}
    ms = AD.ThreadMessenger.group(2)
    ts = [threading.Thread(target=_run_rank, args=(m, 8, 4, 5, 4, 3, 4, out)) for m in ms]
    [t.start() for t in ts]
    [t.join(timeout=60) for t in ts]
    (_hist, val, _e, _s), _q, eps, _lr, batches = out[0]
    assert sum(batches) == 32 and max(batches) <= 3 and len(val) == 6 and eps == 33.0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_rank(rank, world, port, ret):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = {}
        _run_rank(AD.TorchDistMessenger(), 5, 6, 6, 6, 8, 1, out)
        (hist, val, env, state), q, eps, lr, batches = out[rank]
        ret[rank] = (list(map(float, hist)), val, env is None, None if state is None else sorted(state), q.tolist(), eps, lr, batches)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks_train(world):
    import torch.multiprocessing as mp

    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gloo_rank, args=(world, port, ret), nprocs=world, join=True)
        hist, val, env_none, state, q, eps, lr, batches = ret[0]
        assert val == [] and env_none and state is None and q == [[0.0, 5.0]] and eps == 6.0 and lr == 1.0 and sum(batches) == 5
        for r in range(1, world):
            hist, val, env_none, state, *_ = ret[r]
            assert hist == [] and val == [] and not env_none and "states" in state
