"""Generate ``tests/golden/*.npz`` by running the REAL reference (TEST INFRASTRUCTURE).

Run in the build container only (``/root/reference`` does not travel to the GPU box)::

    python oracle/make_golden.py

The reference's own classes are imported unmodified from ``/root/reference/src`` on top
of ``oracle/gym_stub`` (gymnasium is absent here); randomness is injected through the
reference's documented seams: ``algo._rng`` / ``algo._np_rng`` (precedent T-RT:70,
T-MPI:36-37) and ``TicTacToeEnv._np_random`` (rebound by ``reset`` through
``gymnasium.utils.seeding.np_random``, TTT:88-94, which is patched to hand back the shim).
Each fixture stores the inputs, the pre-drawn uniforms and the reference's outputs.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "gym_stub"))
sys.path.insert(0, "/root/reference/src")

from dist_classicrl.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase  # noqa: E402
from dist_classicrl.algorithms.runtime.single_thread_runtime import SingleThreadQLearning  # noqa: E402
from dist_classicrl.environments import tiktaktoe_mod  # noqa: E402
from dist_classicrl.environments.rigged_two_armed_bandit import RiggedTwoArmedBanditEnv  # noqa: E402
from dist_classicrl.schedules.constant_schedule import ConstantSchedule  # noqa: E402
from dist_classicrl.schedules.exponential_schedule import ExponentialSchedule  # noqa: E402
from dist_classicrl.schedules.linear_schedule import LinearSchedule  # noqa: E402
from dist_classicrl.utils import _make_dummy_vec_env  # noqa: E402
from dist_classicrl.wrappers.flatten_multidiscrete_wrapper import (  # noqa: E402
    FlattenMultiDiscreteObservationsWrapper,
)
from gymnasium.vector import SyncVectorEnv  # noqa: E402
from gymnasium.vector.vector_env import AutoresetMode  # noqa: E402

from oracle.envs import T_INIT, HashMDPVec  # noqa: E402
from oracle.rng import SLOT_ENV0, SLOT_ENV1, SLOT_ENV2, draw_uniforms, install_predrawn, pick, u01  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _quantised_table(rng, s, a, dtype, levels=4):
    """Random table with few distinct values so that arg-max ties are common."""
    return (rng.integers(0, levels, size=(s, a)) / levels).astype(dtype)


# ----------------------------------------------------------------------------- select
def gen_select():
    cases = [
        # name, S, A, N, masks?, deterministic, eps, empty_rows
        ("iter_masked_a9", 50, 9, 128, True, False, 0.3, True),  # QLO:714-720 -> choose_masked_action
        ("veciter_masked_a16", 200, 16, 256, True, False, 0.3, True),  # QLO:721-726
        ("iter_nomask_a4", 30, 4, 50, False, False, 0.5, False),  # QLO:701-705
        ("veciter_nomask_a200", 20, 200, 150, False, False, 0.5, False),  # QLO:706-709
        ("vec_nomask_a16", 40, 16, 150, False, False, 0.5, False),  # QLO:710-712 (batch _np_rng)
        ("det_iter_masked_a9", 50, 9, 64, True, True, 0.0, False),  # QLO:671-678
        ("det_vec_masked_a16", 50, 16, 64, True, True, 0.0, False),  # QLO:690-696
        ("det_vec_nomask_a16", 50, 16, 64, False, True, 0.0, False),  # QLO:697-699
        ("eps0_masked_a9", 50, 9, 64, True, False, 0.0, False),
        ("eps1_masked_a16", 50, 16, 64, True, False, 1.0, False),
    ]
    out = {}
    for ci, (name, s, a, n, use_mask, det, eps, empty_rows) in enumerate(cases):
        rng = np.random.default_rng(100 + ci)
        algo = OptimalQLearningBase(s, a, 0.97, seed=0)
        algo.q_table = _quantised_table(rng, s, a, np.float32)
        states = rng.integers(0, s, size=n).astype(np.int32)
        masks = None
        if use_mask:
            masks = (rng.random((n, a)) < 0.6).astype(np.int32)
            masks[np.arange(n), rng.integers(0, a, size=n)] = 1  # at least one legal action
        u = draw_uniforms(7 + ci, 0, 1, n, 2)
        if empty_rows and a <= 10:
            masks[::17] = 0  # all-zero masks -> -1 in the iter variant (QLO:348)
        elif empty_rows:
            # vec variant: an all-zero mask ties every action at -inf on exploit (QLO:467-470);
            # on explore the reference raises IndexError (choice of an empty array), so only
            # exploiting rows get an empty mask here.
            exploit_rows = np.nonzero(u01(u[0, :, 0]) >= eps)[0][::5]
            masks[exploit_rows] = 0
        shim = install_predrawn(algo, u)
        shim.begin_step(0, eps)
        actions = algo.choose_actions(states, eps, deterministic=det, action_masks=masks)
        assert actions.dtype == np.int32
        out[f"{name}__q"] = algo.q_table
        out[f"{name}__states"] = states
        if masks is not None:
            out[f"{name}__masks"] = masks
        out[f"{name}__u"] = u[0]
        out[f"{name}__eps"] = np.float64(eps)
        out[f"{name}__det"] = np.bool_(det)
        out[f"{name}__actions"] = actions
    np.savez_compressed(os.path.join(OUT, "select.npz"), **out)
    print("select.npz:", len(cases), "cases")


# ----------------------------------------------------------------------------- learn
def gen_learn():
    cases = [
        # name, S, A, N, masks?, dtype, p_term
        ("dense_masked_f32", 64, 9, 512, True, np.float32, 0.1),  # N >> S: long same-cell chains
        ("nEqS_masked_f32", 1000, 16, 1000, True, np.float32, 0.05),  # cross-row hazards
        ("nEqS_nomask_f32", 1000, 8, 1000, False, np.float32, 0.05),
        ("sparse_masked_f32", 8192, 8, 256, True, np.float32, 0.05),
        ("dense_masked_f64", 64, 9, 512, True, np.float64, 0.1),
        ("chain_f32", 8, 4, 256, True, np.float32, 0.0),  # s' of agent i == s of agent i-1 etc.
    ]
    out = {}
    for ci, (name, s, a, n, use_mask, dtype, p_term) in enumerate(cases):
        rng = np.random.default_rng(200 + ci)
        algo = OptimalQLearningBase(s, a, 0.97, seed=0)
        algo.q_table = rng.standard_normal((s, a)).astype(dtype)
        q0 = algo.q_table.copy()
        states = rng.integers(0, s, size=n).astype(np.int32)
        actions = rng.integers(0, a, size=n).astype(np.int32)
        rewards = rng.standard_normal(n).astype(np.float32)
        next_states = rng.integers(0, s, size=n).astype(np.int32)
        term = rng.random(n) < p_term
        masks = None
        if use_mask:
            masks = (rng.random((n, a)) < 0.6).astype(np.int32)
            masks[np.arange(n), rng.integers(0, a, size=n)] = 1
        lr = 0.37
        algo.learn(states, actions, rewards, next_states, term, lr, masks)
        out[f"{name}__q0"] = q0
        out[f"{name}__states"] = states
        out[f"{name}__actions"] = actions
        out[f"{name}__rewards"] = rewards
        out[f"{name}__next_states"] = next_states
        out[f"{name}__term"] = term
        if masks is not None:
            out[f"{name}__masks"] = masks
        out[f"{name}__lr"] = np.float64(lr)
        out[f"{name}__gamma"] = np.float64(0.97)
        out[f"{name}__q1"] = algo.q_table
        # accumulate variant (learn_vec) from the same q0
        algo2 = OptimalQLearningBase(s, a, 0.97, seed=0)
        algo2.q_table = q0.copy()
        algo2.learn_vec(states, actions, rewards, next_states, term, lr, masks)
        out[f"{name}__q1_vec"] = algo2.q_table
    np.savez_compressed(os.path.join(OUT, "learn.npz"), **out)
    print("learn.npz:", len(cases), "cases")


# ----------------------------------------------------------------------------- trajectories
class _StepClock:
    """Shared step counter: advanced by the ``choose_actions`` hook (one call per vector step)."""

    def __init__(self):
        self.t = T_INIT


class _TTTEnvRng:
    """``TicTacToeEnv._np_random`` shim: dispatches on the call site."""

    def __init__(self, clock, u_steps, u_init, i):
        self.clock, self.u_steps, self.u_init, self.i = clock, u_steps, u_init, i

    def choice(self, seq):
        u = self.u_init if self.clock.t == T_INIT else self.u_steps[self.clock.t]
        if isinstance(seq, range):  # TTT:106 choice(range(9))
            slot = SLOT_ENV2
        elif len(seq) == 2 and seq[0] is True:  # TTT:98 choice([True, False])
            slot = SLOT_ENV1
        else:  # TTT:185 choice(valid_moves)
            slot = SLOT_ENV0
        return seq[pick(int(u[self.i, slot]), len(seq))]


class _Recorder:
    """Env proxy recording what ``run_single_step`` sees (BRT:210)."""

    def __init__(self, env, f32_rewards=False):
        self._env = env
        self.num_envs = env.num_envs
        self.f32_rewards = f32_rewards
        self.trace = {"actions": [], "rewards": [], "terminated": [], "obs": []}

    def reset(self, **kw):
        return self._env.reset(**kw)

    def step(self, actions):
        out = self._env.step(actions)
        if self.f32_rewards:  # DistClassicRLEnv contract (ENV:53): float32 rewards -> pure fp32 TD update
            out = (out[0], np.asarray(out[1], dtype=np.float32), *out[2:])
        obs = out[0]["observation"] if isinstance(out[0], dict) else out[0]
        self.trace["actions"].append(np.asarray(actions).copy())
        self.trace["rewards"].append(np.asarray(out[1], dtype=np.float32).copy())
        self.trace["terminated"].append(np.asarray(out[2]).copy())
        self.trace["obs"].append(np.asarray(obs).copy())
        return out

    def stacked(self):
        return {k: np.stack(v) for k, v in self.trace.items()}


def _hook_clock(algo, shim, clock, eps_sched):
    real = algo.choose_actions
    counter = {"t": 0}

    def hooked(*args, **kwargs):
        clock.t = counter["t"]
        shim.begin_step(counter["t"], eps_sched.get_value())
        counter["t"] += 1
        return real(*args, **kwargs)

    algo.choose_actions = hooked


def _run_reference(algo, env, steps, lr_sched, eps_sched, uniforms, clock, f32_rewards=False):
    shim = install_predrawn(algo, uniforms)
    _hook_clock(algo, shim, clock, eps_sched)
    rt = SingleThreadQLearning(algo, lr_sched, eps_sched)
    rec = _Recorder(env, f32_rewards)
    clock.t = T_INIT
    try:
        avg, history, _, sd = rt.run_steps(steps, rec, None)
    except ZeroDivisionError:  # STR:67 when no episode finished
        history, sd = [], None
    return history, rec.stacked(), sd


def _ttt_run(n, steps, seed, f32_rewards):
    u_steps = draw_uniforms(seed, 0, steps, n, 5)
    u_init = draw_uniforms(seed, T_INIT, 1, n, 5)[0]
    clock = _StepClock()
    shims = {}
    orig_np_random = tiktaktoe_mod.np_random
    tiktaktoe_mod.np_random = lambda s, _sh=shims: (_sh[s], s)  # TTT:92 seam
    try:
        env = SyncVectorEnv(
            [lambda: FlattenMultiDiscreteObservationsWrapper(tiktaktoe_mod.TicTacToeEnv()) for _ in range(n)],
            autoreset_mode=AutoresetMode.SAME_STEP,
        )
        for i, e in enumerate(env.envs):
            shims[i] = _TTTEnvRng(clock, u_steps, u_init, i)
            e.env._np_random_seed = i  # makes un-seeded reset() call np_random(i) (TTT:89-92)
        algo = OptimalQLearningBase(19683, 9, 0.99, seed=0)
        algo.q_table = algo.q_table.astype(np.float32)
        # TPB:157-166 schedules (slower decay for N=1 so that exploration lasts)
        lr = ExponentialSchedule(0.1, 1e-5, 0.995 if n > 1 else 0.9995)
        eps = ExponentialSchedule(1.0, 0.01, 0.995 if n > 1 else 0.9995)
        history, trace, sd = _run_reference(algo, env, steps, lr, eps, u_steps, clock, f32_rewards)
    finally:
        tiktaktoe_mod.np_random = orig_np_random
    return u_steps, u_init, algo.q_table, history, trace, lr, eps


def gen_ttt():
    """TicTacToe + flatten wrapper + SyncVectorEnv(SAME_STEP) through SingleThreadQLearning.run_steps.

    Two runs per size: ``f32`` casts the env rewards to float32 (the DistClassicRLEnv
    contract, ENV:53) so the TD update is pure fp32 -- the bit-exact target of the CUDA
    engine; ``f64r`` leaves SyncVectorEnv's float64 rewards alone (reference as-is: the
    update is then evaluated in fp64 and rounded on store) -- the <=1e-6 target.
    """
    out = {}
    for name, n, steps, seed in [("n1", 1, 400, 11), ("n16", 16, 300, 12), ("n128", 128, 120, 13)]:
        u_steps, u_init, q, history, trace, lr, eps = _ttt_run(n, steps, seed, True)
        _, _, q64, history64, trace64, _, _ = _ttt_run(n, steps, seed, False)
        same = all(np.array_equal(trace[k], trace64[k]) for k in trace)
        out[f"{name}__u_steps"] = u_steps
        out[f"{name}__u_init"] = u_init
        out[f"{name}__q"] = q
        out[f"{name}__q_f64r"] = q64
        out[f"{name}__f64r_same_trajectory"] = np.bool_(same)
        out[f"{name}__history"] = np.asarray(history, dtype=np.float32)
        out[f"{name}__decay"] = np.float64(lr.decay_rate)
        out[f"{name}__lr_end"] = np.float64(lr.get_value())
        out[f"{name}__eps_end"] = np.float64(eps.get_value())
        for k, v in trace.items():
            out[f"{name}__{k}"] = v
        print(f"ttt {name}: {len(history)} episodes, |q|max={np.abs(q).max():.4f}, f64-reward run: same trajectory={same}, "
              f"max|dq|={np.abs(q.astype(np.float64) - q64).max():.3e}")
    np.savez_compressed(os.path.join(OUT, "ttt_traj.npz"), **out)


class _MDPGym:
    """``DistClassicRLEnv``-shaped adapter of the NumPy hash MDP (float32 rewards, ENV:53)."""

    def __init__(self, mdp, clock, u_steps, u_init):
        self.mdp, self.clock, self.u_steps, self.u_init = mdp, clock, u_steps, u_init
        self.num_envs = self.num_agents = mdp.num_envs

    def reset(self, seed=None, options=None):
        return self.mdp.reset(self.u_init)

    def step(self, actions):
        return self.mdp.step(actions, self.u_steps[self.clock.t])


def gen_mdp():
    out = {}
    for name, s, a, n, steps, seed in [("s1000_a16", 1000, 16, 512, 40, 21), ("s200_a8", 200, 8, 300, 40, 22)]:
        u_steps = draw_uniforms(seed, 0, steps, n, 4)
        u_init = draw_uniforms(seed, T_INIT, 1, n, 4)[0]
        clock = _StepClock()
        env = _MDPGym(HashMDPVec(n, s, a, seed=seed, p_term=0.05), clock, u_steps, u_init)
        algo = OptimalQLearningBase(s, a, 0.99, seed=0)
        algo.q_table = algo.q_table.astype(np.float32)
        lr = ExponentialSchedule(0.5, 1e-3, 0.9999)
        eps = ExponentialSchedule(1.0, 0.05, 0.9995)
        history, trace, sd = _run_reference(algo, env, steps, lr, eps, u_steps, clock)
        out[f"{name}__u_steps"] = u_steps
        out[f"{name}__u_init"] = u_init
        out[f"{name}__cfg"] = np.asarray([s, a, n, steps, seed], dtype=np.int64)
        out[f"{name}__q"] = algo.q_table
        out[f"{name}__history"] = np.asarray(history, dtype=np.float32)
        out[f"{name}__lr_end"] = np.float64(lr.get_value())
        out[f"{name}__eps_end"] = np.float64(eps.get_value())
        for k, v in trace.items():
            out[f"{name}__{k}"] = v
        print(f"mdp {name}: {len(history)} episodes, |q|max={np.abs(algo.q_table).max():.4f}")
    np.savez_compressed(os.path.join(OUT, "mdp_traj.npz"), **out)


def gen_bandit():
    """T-RT:77-98 shape: bandit(episode_len=5) in DummyVecWrapper, 3 envs, 23 steps."""
    n, steps = 3, 23
    u_steps = draw_uniforms(31, 0, steps, n, 2)
    clock = _StepClock()
    env = _make_dummy_vec_env(n, RiggedTwoArmedBanditEnv, {"episode_len": 5})
    algo = OptimalQLearningBase(1, 2, 0.9, seed=0)
    algo.q_table = algo.q_table.astype(np.float32)
    lr = ConstantSchedule(0.25)
    eps = LinearSchedule(0.9, -0.01)
    history, trace, sd = _run_reference(algo, env, steps, lr, eps, u_steps, clock)
    out = {"u_steps": u_steps, "q": algo.q_table, "history": np.asarray(history, dtype=np.float32),
           "eps_end": np.float64(eps.get_value())}
    out.update(trace)
    np.savez_compressed(os.path.join(OUT, "bandit_traj.npz"), **out)
    print("bandit:", len(history), "episodes", algo.q_table)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_select()
    gen_learn()
    gen_ttt()
    gen_mdp()
    gen_bandit()
