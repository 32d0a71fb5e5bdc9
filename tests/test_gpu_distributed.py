"""Multi-GPU modes on ONE device: G virtual ranks (threads, LoopbackTransport) share the GPU.

Sharded table (SURVEY 8e / BASELINE config 4): table, agent states and running returns after K vector steps must be
identical -- bit for bit -- to the single-GPU fused loop AND to the C oracle on the same seeds, for any G.  The G ranks
run side by side in ONE cooperative launch of the peer-memory kernel (csrc/qe_shard.cuh), a group of CTAs per rank.
Replicated table (config 5): the merged table must equal the NumPy restatement of the delta rule applied to the
oracle's per-rank tables (exact at G = 2: a two-term sum is order-free)."""
import ctypes as C
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import c_oracle as co  # noqa: E402
from oracle import rng as orng  # noqa: E402
from oracle.envs import T_INIT  # noqa: E402

EPS, LR, GAMMA, P_TERM = 0.1, 0.1, 0.99, 0.05


@pytest.fixture(scope="module")
def capi():
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    from dist_classicrl_b200 import capi as m

    m.lib()
    return m


def _oracle_run(S, A, N, steps, seed, env_seed, table_seed, agent0=0, q0=None):
    tt = int(math.ceil(P_TERM * 2.0**32))
    states, masks = co.mdp_reset(orng.draw_uniforms(seed, T_INIT, 1, N, 4, agent0=agent0)[0], S, A, env_seed)
    q = q0.copy() if q0 is not None else _random_table(S, A, table_seed)
    th = np.full(steps, orng.explore_threshold(EPS), dtype=np.uint64)
    lr = np.full(steps, LR, dtype=np.float32)
    rew = np.zeros(N, dtype=np.float32)
    res = co.run(co.ENV_MDP, q, None, states, masks, num_states=S, env_seed=env_seed, term_thresh=tt, uniforms=None, slots=4,
                 stream_seed=seed, t0=0, agent0=agent0, steps=steps, eps_thresh=th, lr=lr, gamma=GAMMA, empty_all=A > 10,
                 agent_rewards=rew)
    assert res["rc"] == 0
    return q, states, rew


def _random_table(S, A, table_seed):
    """qe_table_fill_random restated: (fmix32((s*A+a) ^ seed*GOLD) >> 8) * 2^-24."""
    x = (np.arange(S * A, dtype=np.uint64) ^ np.uint64((table_seed * 0x9E3779B9) & 0xFFFFFFFF)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return ((x >> np.uint64(8)).astype(np.float32) * np.float32(2.0**-24)).reshape(S, A)


@pytest.mark.parametrize("world,S,A,N,steps", [(2, 5000, 8, 6000, 10), (3, 1000, 16, 4000, 8), (4, 200_000, 8, 30_000, 6),
                                               (2, 64, 4, 512, 6), (8, 30_000, 8, 100_000, 12), (8, 7, 20, 3000, 5),
                                               (5, 1_000_000, 16, 200_001, 9)])
def test_sharded_table_matches_oracle(capi, world, S, A, N, steps):
    from dist_classicrl_b200 import distributed as D
    from dist_classicrl_b200.schedules import ConstantSchedule

    seed, env_seed, table_seed = 7, 3, 1
    q_o, st_o, rew_o = _oracle_run(S, A, N, steps, seed, env_seed, table_seed)

    def body(tp):
        torch.cuda.set_device(0)
        sh = D.ShardedQLearning(S, A, GAMMA, N, tp, env_seed=env_seed, p_term=P_TERM, seed=seed, device=0)
        sh.fill_random(table_seed)
        sh.reset()
        sh.run_steps(steps, ConstantSchedule(EPS), ConstantSchedule(LR))
        table = sh.gather_table()
        states, rets = sh.gather_agents()
        cs = sh.table_checksum()
        sh.close()
        return table, states, rets, cs

    out = D.run_loopback(world, body)
    table, states, rets, cs = out[0]
    assert np.array_equal(states, st_o)
    assert np.array_equal(table, q_o), f"max |dq| = {np.abs(table - q_o).max()}"
    assert np.array_equal(rets, rew_o)
    for other in out[1:]:
        assert np.array_equal(other[0], table) and other[3] == cs


@pytest.mark.parametrize("stepwise", [False, True])
def test_replicated_table_delta_rule(capi, stepwise):
    from dist_classicrl_b200 import distributed as D
    from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase
    from dist_classicrl_b200.algorithms.runtime import SingleThreadQLearning
    from dist_classicrl_b200.environments import HashMDPVecEnv
    from dist_classicrl_b200.schedules import ConstantSchedule

    world, S, A, n_local, sync_every, chunks = 2, 3000, 8, 2048, 4, 3
    seed, env_seed, table_seed = 11, 2, 1

    def body(tp):
        torch.cuda.set_device(0)
        algo = OptimalQLearningBase(S, A, GAMMA, seed=seed, device=0)
        algo.fill_random(table_seed)
        env = HashMDPVecEnv(n_local, S, A, env_seed=env_seed, p_term=P_TERM, seed=seed, device=0, output="torch")
        env.agent0 = tp.rank * n_local
        env.attach(algo)
        rt = SingleThreadQLearning(algo, ConstantSchedule(LR), ConstantSchedule(EPS))
        rt.history_mode = "summary"
        if stepwise:  # one vector step per call, the merge period carried across the calls: same merges, same result
            rep = D.ReplicatedQLearning(rt, tp, sync_every=sync_every, carry_over=True)
            state = None
            for _ in range(sync_every * chunks):
                _m, _h, env, state = rep.run_steps(1, env, state)
        else:
            rep = D.ReplicatedQLearning(rt, tp, sync_every=sync_every)
            rep.run_steps(sync_every * chunks, env)
        return np.array(algo.q_table, copy=True), env.states.cpu().numpy(), rep.syncs

    out = D.run_loopback(world, body)
    # restatement: every rank runs `sync_every` oracle steps from the common base, then base += sum of deltas (fp32)
    tt = int(math.ceil(P_TERM * 2.0**32))
    base = _random_table(S, A, table_seed)
    st = [co.mdp_reset(orng.draw_uniforms(seed, T_INIT, 1, n_local, 4, agent0=r * n_local)[0], S, A, env_seed) for r in range(world)]
    th = np.full(sync_every, orng.explore_threshold(EPS), dtype=np.uint64)
    lr = np.full(sync_every, LR, dtype=np.float32)
    for c in range(chunks):
        deltas = []
        for r in range(world):
            q = base.copy()
            res = co.run(co.ENV_MDP, q, None, st[r][0], st[r][1], num_states=S, env_seed=env_seed, term_thresh=tt, uniforms=None,
                         slots=4, stream_seed=seed, t0=c * sync_every, agent0=r * n_local, steps=sync_every, eps_thresh=th, lr=lr,
                         gamma=GAMMA, empty_all=A > 10)
            assert res["rc"] == 0
            deltas.append(q - base)
        base = base + (deltas[0] + deltas[1])
    for r in range(world):
        assert out[r][2] == chunks
        assert np.array_equal(out[r][1], st[r][0])
        assert np.array_equal(out[r][0], base), f"rank {r}: max |dq| = {np.abs(out[r][0] - base).max()}"
