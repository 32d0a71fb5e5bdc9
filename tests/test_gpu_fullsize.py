"""Parity at BASELINE.json's full sizes.

Config 3 (1M states x 16 actions, 2^20 agents) is small enough for the C oracle: bit-exact table, states and
returns after 12 vector steps (well into the clustered regime: half of the agents share a row with more than four
others).  Config 4's table (100M x 8, 2^22 agents) is checked on one GPU against the oracle on the visited rows,
plus the size-independent properties of the fused loop: chunking (K steps in one launch == K launches of one
step) and run-to-run determinism."""
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
TT = int(math.ceil(P_TERM * 2.0**32))
_CACHE = {}


@pytest.fixture(scope="module")
def capi():
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    from dist_classicrl_b200 import capi as m

    m.lib()
    return m


class Run:
    """One engine + hash-MDP agents driven through qe_fused_steps."""

    def __init__(self, capi, S, A, N, seed, table_seed):
        self.capi, self.lib, self.S, self.A, self.N, self.seed = capi, capi.lib(), S, A, N, seed
        self.h = C.c_void_p()
        capi.check(self.lib.qe_create(S, A, GAMMA, 0, C.byref(self.h)))
        capi.check(self.lib.qe_table_fill_random(self.h, table_seed, None))
        dev = torch.device("cuda:0")
        self.states = torch.empty(N, dtype=torch.int32, device=dev)
        self.scratch = torch.empty_like(self.states)
        self.ep = torch.zeros(N, dtype=torch.float32, device=dev)
        capi.check(self.lib.qe_mdp_reset(self.states.data_ptr(), None, S, A, seed, None, 4, seed, T_INIT, 0, N, None))
        self.ag = capi.QeAgents(capi.QE_ENV_MDP, N, self.states.data_ptr(), self.scratch.data_ptr(), None, self.ep.data_ptr(), seed, 0, TT)
        self.t = 0

    def steps(self, k):
        th = np.full(k, orng.explore_threshold(EPS), dtype=np.uint64)
        lr = np.full(k, LR, dtype=np.float32)
        run = self.capi.QeRun()
        run.steps = k
        run.explore_thresholds_host, run.learning_rates_host = th.ctypes.data_as(C.c_void_p), lr.ctypes.data_as(C.c_void_p)
        run.slots = 4
        run.stream_seed = run.env_stream_seed = self.seed
        run.t0 = run.env_t0 = self.t
        run.use_masks, run.empty_all = 1, int(self.A > 10)
        self.capi.check(self.lib.qe_fused_steps(self.h, C.byref(self.ag), C.byref(run), None))
        self.capi.check(self.lib.qe_sync(self.h, None))
        self.t += k

    def rows(self, states):
        states = np.ascontiguousarray(states, dtype=np.int32)
        out = np.empty((states.shape[0], self.A), dtype=np.float32)
        self.capi.check(self.lib.qe_gather_rows_host(self.h, states.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), states.shape[0]))
        return out

    def table(self):
        q = np.empty((self.S, self.A), dtype=np.float32)
        self.capi.check(self.lib.qe_table_download_host(self.h, q.ctypes.data_as(C.c_void_p)))
        return q

    def close(self):
        self.lib.qe_destroy(self.h)


def _random_table(S, A, table_seed):
    x = (np.arange(S * A, dtype=np.uint64) ^ np.uint64((table_seed * 0x9E3779B9) & 0xFFFFFFFF)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return ((x >> np.uint64(8)).astype(np.float32) * np.float32(2.0**-24)).reshape(S, A)


def _oracle(S, A, N, steps, seed, table_seed):
    st, mk = co.mdp_reset(orng.draw_uniforms(seed, T_INIT, 1, N, 4)[0], S, A, seed)
    q = _random_table(S, A, table_seed)
    rew = np.zeros(N, dtype=np.float32)
    visited = [st.copy()]
    th = np.full(1, orng.explore_threshold(EPS), dtype=np.uint64)
    lr = np.full(1, LR, dtype=np.float32)
    for t in range(steps):
        res = co.run(co.ENV_MDP, q, None, st, mk, num_states=S, env_seed=seed, term_thresh=TT, uniforms=None, slots=4, stream_seed=seed,
                     t0=t, steps=1, eps_thresh=th, lr=lr, gamma=GAMMA, empty_all=A > 10, agent_rewards=rew)
        assert res["rc"] == 0
        visited.append(st.copy())
    return q, st, rew, np.unique(np.concatenate(visited))


def test_config3_full_size_bit_exact_vs_oracle(capi):
    S, A, N, steps, seed = 1_000_000, 16, 1 << 20, 12, 0
    q_o, st_o, rew_o, _ = _oracle(S, A, N, steps, seed, 1)
    cnt = np.bincount(st_o, minlength=S)
    assert np.mean(cnt[st_o] > 4) > 0.3, "the run must reach the clustered regime"
    r = Run(capi, S, A, N, seed, 1)
    try:
        r.steps(5)
        r.steps(7)
        assert np.array_equal(r.states.cpu().numpy(), st_o)
        assert np.array_equal(r.ep.cpu().numpy(), rew_o)
        assert np.array_equal(r.table(), q_o)
    finally:
        r.close()


def test_config4_table_size_visited_rows_vs_oracle(capi):
    S, A, N, steps, seed = 100_000_000, 8, 1 << 22, 3, 0
    q_o, st_o, rew_o, visited = _oracle(S, A, N, steps, seed, 1)
    r = Run(capi, S, A, N, seed, 1)
    try:
        r.steps(steps)
        assert np.array_equal(r.states.cpu().numpy(), st_o)
        assert np.array_equal(r.ep.cpu().numpy(), rew_o)
        assert np.array_equal(r.rows(visited), q_o[visited])  # every row any agent ever stood on
        probe = np.arange(0, S, 9973, dtype=np.int32)          # and a stride of untouched ones
        assert np.array_equal(r.rows(probe), q_o[probe])
    finally:
        r.close()


def test_config3_chunking_and_determinism(capi):
    S, A, N, seed = 1_000_000, 16, 1 << 20, 3
    a, b, c = (Run(capi, S, A, N, seed, 1) for _ in range(3))
    try:
        a.steps(8)
        for _ in range(8):
            b.steps(1)
        c.steps(3)
        c.steps(5)
        sa, ea, qa = a.states.cpu().numpy(), a.ep.cpu().numpy(), a.table()
        for other in (b, c):
            assert np.array_equal(other.states.cpu().numpy(), sa)
            assert np.array_equal(other.ep.cpu().numpy(), ea)
            assert np.array_equal(other.table(), qa)
    finally:
        for x in (a, b, c):
            x.close()


@pytest.mark.parametrize("form", ["0", "1", "3", "5"])
@pytest.mark.parametrize("S,A,N,steps", [(2000, 16, 50_000, 24), (97, 5, 4096, 16), (300_000, 8, 200_000, 10), (3, 20, 1000, 6), (70_000, 32, 33, 40), (3_000_000, 4, 150_000, 8)])
def test_all_forms_of_the_fused_update_match_the_oracle(capi, monkeypatch, form, S, A, N, steps):
    """QE_FORM pins the form of the TD update (0 = writer lists, 1 = per-step sort, 3 = target pipeline, 5 = one-pass form); all must
    reproduce the oracle bit for bit, also when hundreds of agents herd on one row (or all of them on three rows)."""
    monkeypatch.setenv("QE_FORM", form)
    seed = 9
    q_o, st_o, rew_o, _ = _oracle(S, A, N, steps, seed, 1)
    assert np.bincount(st_o, minlength=S).max() > (40 if N > 10 * S else 0)
    r = Run(capi, S, A, N, seed, 1)
    try:
        r.steps(steps // 2)
        r.steps(steps - steps // 2)
        assert capi.lib().qe_fused_form(r.h) == (4 if (form in ("3", "5") and N <= 256) else int(form))  # batches of <= 256 agents: the one-CTA loop
        assert np.array_equal(r.states.cpu().numpy(), st_o)
        assert np.array_equal(r.ep.cpu().numpy(), rew_o)
        assert np.array_equal(r.table(), q_o)
    finally:
        r.close()


@pytest.mark.parametrize("flags", ["1", "5"])
@pytest.mark.parametrize("S,A,N,steps", [(2000, 16, 50_000, 12), (300_000, 8, 200_000, 8), (3, 20, 1000, 6), (3_000_000, 4, 150_000, 6), (1_000_000, 16, 1 << 18, 6)])
def test_one_pass_form_every_bucket_sort_path_matches_the_oracle(capi, monkeypatch, flags, S, A, N, steps):
    """The next step's order is built by whichever bucket sort fits: teams of two warps, one warp for tiny buckets, the whole
    block in shared memory, counting passes through global memory for what does not fit there.  QE_FLOW_FLAGS = 1 takes the
    teams out, 5 also limits the shared-memory path to 64 keys, so that ordinary buckets exercise the other paths (one low
    digit, two low digits with the copy back, no low bits at all); the results must not change."""
    monkeypatch.setenv("QE_FORM", "5")
    monkeypatch.setenv("QE_FLOW_FLAGS", flags)
    seed = 11
    q_o, st_o, rew_o, _ = _oracle(S, A, N, steps, seed, 1)
    r = Run(capi, S, A, N, seed, 1)
    try:
        r.steps(steps // 2)
        r.steps(steps - steps // 2)
        assert capi.lib().qe_fused_form(r.h) == 5
        assert np.array_equal(r.states.cpu().numpy(), st_o)
        assert np.array_equal(r.ep.cpu().numpy(), rew_o)
        assert np.array_equal(r.table(), q_o)
    finally:
        r.close()


@pytest.mark.parametrize("S,A,N,steps", [(19, 4, 1, 60), (500, 9, 128, 200), (40, 32, 256, 50), (5000, 16, 255, 64), (3, 8, 200, 30)])
def test_small_batches_one_cta_loop_matches_the_oracle(capi, monkeypatch, S, A, N, steps):
    """Batches of at most 256 agents run in one CTA (csrc/qe_small.cuh): same results as the oracle, whatever the crowding."""
    monkeypatch.delenv("QE_FORM", raising=False)
    seed = 5
    q_o, st_o, rew_o, _ = _oracle(S, A, N, steps, seed, 1)
    r = Run(capi, S, A, N, seed, 1)
    try:
        r.steps(steps // 3)
        r.steps(steps - steps // 3)
        assert capi.lib().qe_fused_form(r.h) == 4
        assert np.array_equal(r.states.cpu().numpy(), st_o)
        assert np.array_equal(r.ep.cpu().numpy(), rew_o)
        assert np.array_equal(r.table(), q_o)
    finally:
        r.close()


def test_automatic_form_selection_switches_forms_and_stays_exact(capi, monkeypatch):
    """Strategy 2 (QE_FORM=2; round 1's default): the engine's first four launches are {writer lists cold, timed, sort cold,
    timed}, then it keeps the faster form and times both again in two adjacent launches every 12 launches.  Whatever
    it picks, the results are the oracle's; qe_set_fused_form pins and releases the choice."""
    monkeypatch.delenv("QE_SORTED", raising=False)
    monkeypatch.setenv("QE_FORM", "2")
    S, A, N, launches, k, seed = 4000, 16, 60_000, 32, 1, 4
    q_o, st_o, rew_o, _ = _oracle(S, A, N, launches * k + 12, seed, 1)
    r = Run(capi, S, A, N, seed, 1)
    try:
        forms = []
        for _ in range(launches):
            r.steps(k)
            forms.append(int(capi.lib().qe_fused_form(r.h)))
        assert forms[:4] == [0, 0, 1, 1]
        steady = forms[4]
        assert forms[4:15] == [steady] * 11            # launches 5..15 use the pick
        assert forms[15:17] == [steady ^ 1, steady]    # the probe: the other form, then the current one, adjacent
        assert set(forms[17:28]) <= {0, 1} and len(set(forms[17:28])) == 1
        capi.check(capi.lib().qe_set_fused_form(r.h, 0))
        r.steps(2)
        assert capi.lib().qe_fused_form(r.h) == 0
        capi.check(capi.lib().qe_set_fused_form(r.h, 1))
        r.steps(2)
        assert capi.lib().qe_fused_form(r.h) == 1
        capi.check(capi.lib().qe_set_fused_form(r.h, 3))
        r.steps(2)
        assert capi.lib().qe_fused_form(r.h) == 3
        capi.check(capi.lib().qe_set_fused_form(r.h, 5))
        r.steps(1)
        r.steps(3)
        assert capi.lib().qe_fused_form(r.h) == 5
        capi.check(capi.lib().qe_set_fused_form(r.h, 3))
        r.steps(2)
        assert capi.lib().qe_fused_form(r.h) == 3
        with pytest.raises(ValueError):
            capi.check(capi.lib().qe_set_fused_form(r.h, 4))
        assert np.array_equal(r.states.cpu().numpy(), st_o)
        assert np.array_equal(r.ep.cpu().numpy(), rew_o)
        assert np.array_equal(r.table(), q_o)
    finally:
        r.close()


@pytest.mark.parametrize("form", ["5", "3", "1", "0"])
def test_config3_long_run_every_form_vs_oracle(capi, monkeypatch, form):
    """2^20 agents x 72 vector steps (the bench window and beyond; the agents herd: rows with dozens of writers),
    every form of the exact update pinned, bit-exact against the C oracle."""
    monkeypatch.setenv("QE_FORM", form)
    S, A, N, steps, seed = 1_000_000, 16, 1 << 20, 72, 0
    key = ("c3long", steps)
    if key not in _CACHE:
        _CACHE[key] = _oracle(S, A, N, steps, seed, 1)
    q_o, st_o, rew_o, _ = _CACHE[key]
    r = Run(capi, S, A, N, seed, 1)
    try:
        for k in (8, 8, 8, 8, 8, 8, 8, 8, 1, 7):
            r.steps(k)
        assert capi.lib().qe_fused_form(r.h) == int(form)
        assert np.array_equal(r.states.cpu().numpy(), st_o)
        assert np.array_equal(r.ep.cpu().numpy(), rew_o)
        assert np.array_equal(r.table(), q_o)
    finally:
        r.close()
