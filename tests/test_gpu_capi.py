"""GPU parity tests at the C-ABI level: CUDA kernels vs the oracle (golden fixtures + C/NumPy restatement).

Bit-exact: action indices, env states, rewards, done flags AND fp32 Q-values (the CUDA path evaluates the
reference's fp32 formula op by op, so the tolerance for Q is zero here; BASELINE's bar is 1e-6 relative).
"""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import c_oracle as co  # noqa: E402
from oracle import rng as orng  # noqa: E402
from oracle.envs import T_INIT  # noqa: E402
from oracle.runtime import Exponential  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
SELECT = np.load(os.path.join(GOLDEN, "select.npz"))
LEARN = np.load(os.path.join(GOLDEN, "learn.npz"))
TTT = np.load(os.path.join(GOLDEN, "ttt_traj.npz"))
MDP = np.load(os.path.join(GOLDEN, "mdp_traj.npz"))


def _cases(npz):
    return sorted({k.split("__")[0] for k in npz.files})


@pytest.fixture(scope="module")
def capi():
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    from dist_classicrl_b200 import capi as m

    m.lib()
    return m


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())


class Engine:
    def __init__(self, capi, s, a, gamma):
        self.capi, self.S, self.A = capi, s, a
        self.h = C.c_void_p()
        capi.check(capi.lib().qe_create(s, a, gamma, 0, C.byref(self.h)))

    def upload(self, q):
        q = np.ascontiguousarray(q, dtype=np.float32)
        self.capi.check(self.capi.lib().qe_table_upload_host(self.h, _p(q)))

    def download(self):
        q = np.empty((self.S, self.A), dtype=np.float32)
        self.capi.check(self.capi.lib().qe_table_download_host(self.h, _p(q)))
        return q

    def close(self):
        self.capi.lib().qe_destroy(self.h)


def _bytes_or_none(masks):
    return None if masks is None else np.ascontiguousarray(np.asarray(masks) != 0, dtype=np.uint8)


@pytest.mark.parametrize("name", _cases(SELECT))
def test_select_host_matches_reference(capi, name):
    g = lambda k: SELECT[f"{name}__{k}"]  # noqa: E731
    masks = g("masks") if f"{name}__masks" in SELECT.files else None
    q = g("q")
    s, a = q.shape
    e = Engine(capi, s, a, 0.97)
    try:
        e.upload(q)
        states = np.ascontiguousarray(g("states"), dtype=np.int32)
        u = np.ascontiguousarray(g("u"), dtype=np.uint32)
        out = np.empty(states.shape[0], dtype=np.int32)
        det = bool(g("det"))
        empty_all = int(a > 10)
        bits = co.masks_to_bits(masks) if a <= 32 else None
        mbytes = _bytes_or_none(masks) if a > 32 else None
        capi.check(capi.lib().qe_select_host(e.h, _p(states), _p(bits), _p(mbytes), _p(u), u.shape[1], 0, 0, 0,
                                             orng.explore_threshold(float(g("eps"))), int(det), empty_all, _p(out),
                                             states.shape[0]))
        np.testing.assert_array_equal(out, g("actions"))
        if masks is not None and a <= 32:  # byte-mask (generic) kernel must agree as well
            out2 = np.empty_like(out)
            capi.check(capi.lib().qe_select_host(e.h, _p(states), None, _p(_bytes_or_none(masks)), _p(u), u.shape[1], 0, 0, 0,
                                                 orng.explore_threshold(float(g("eps"))), int(det), empty_all, _p(out2),
                                                 states.shape[0]))
            np.testing.assert_array_equal(out2, g("actions"))
    finally:
        e.close()


@pytest.mark.parametrize("name", [c for c in _cases(LEARN) if LEARN[f"{c}__q0"].dtype == np.float32])
@pytest.mark.parametrize("generic", [False, True])
def test_learn_host_matches_reference_bit_exact(capi, name, generic):
    g = lambda k: LEARN[f"{name}__{k}"]  # noqa: E731
    masks = g("masks") if f"{name}__masks" in LEARN.files else None
    q0 = g("q0")
    s, a = q0.shape
    e = Engine(capi, s, a, float(g("gamma")))
    try:
        e.upload(q0)
        n = g("states").shape[0]
        arr = lambda k, dt: np.ascontiguousarray(g(k), dtype=dt)  # noqa: E731
        bits = None if (masks is None or generic) else co.masks_to_bits(masks)
        mbytes = _bytes_or_none(masks) if generic else None
        if generic and masks is None:
            pytest.skip("generic path is selected by byte masks")
        capi.check(capi.lib().qe_learn_host(e.h, _p(arr("states", np.int32)), _p(arr("actions", np.int32)),
                                            _p(arr("rewards", np.float32)), _p(arr("next_states", np.int32)),
                                            _p(arr("term", np.uint8)), _p(bits), _p(mbytes), np.float32(g("lr")), n,
                                            capi.QE_LEARN_SEQUENTIAL))
        np.testing.assert_array_equal(e.download(), g("q1"))
        # accumulate mode (learn_vec): atomics, order-dependent rounding -> tolerance
        e.upload(q0)
        capi.check(capi.lib().qe_learn_host(e.h, _p(arr("states", np.int32)), _p(arr("actions", np.int32)),
                                            _p(arr("rewards", np.float32)), _p(arr("next_states", np.int32)),
                                            _p(arr("term", np.uint8)), _p(bits), _p(mbytes), np.float32(g("lr")), n,
                                            capi.QE_LEARN_ACCUMULATE))
        np.testing.assert_allclose(e.download(), g("q1_vec"), rtol=2e-5, atol=2e-5)
    finally:
        e.close()


def _schedule_arrays(lr, eps, steps, n):
    th, lrs = [], []
    for _ in range(steps):
        th.append(orng.explore_threshold(eps.get_value()))
        lrs.append(np.float32(lr.get_value()))
        lr.update(n)
        eps.update(n)
    return np.asarray(th, dtype=np.uint64), np.asarray(lrs, dtype=np.float32)


def _fused(capi, e, env_kind, n, states, env_words, ep_ret, *, steps, th, lrs, uniforms=None, slots=4, seed=0, t0=0,
           env_seed=0, term_thresh=0, episode_len=0, empty_all=0, use_masks=1, record=True, chunk=None):
    """Run qe_fused_steps (optionally in chunks of `chunk` steps) and return traces as numpy arrays."""
    dev = torch.device("cuda:0")
    scratch = torch.empty_like(states)
    tr = {}
    if record:
        tr = {"actions": torch.empty((steps, n), dtype=torch.int32, device=dev),
              "rewards": torch.empty((steps, n), dtype=torch.float32, device=dev),
              "terminated": torch.empty((steps, n), dtype=torch.uint8, device=dev),
              "obs": torch.empty((steps, n), dtype=torch.int32, device=dev),
              "episode_returns": torch.empty((steps, n), dtype=torch.float32, device=dev)}
    ep_sum = torch.zeros(1, dtype=torch.float64, device=dev)
    ep_cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    ag = capi.QeAgents(env_kind, n, states.data_ptr(), scratch.data_ptr(), env_words.data_ptr() if env_words is not None else None,
                       ep_ret.data_ptr(), env_seed, episode_len, term_thresh)
    # uint32 payload carried in an int32 tensor
    d_u = None if uniforms is None else torch.from_numpy(np.ascontiguousarray(uniforms).view(np.int32)).to(dev)
    chunk = chunk or steps
    k = 0
    while k < steps:
        c = min(chunk, steps - k)
        run = capi.QeRun()
        run.steps = c
        th_c, lr_c = np.ascontiguousarray(th[k:k + c]), np.ascontiguousarray(lrs[k:k + c])
        run.explore_thresholds_host, run.learning_rates_host = _p(th_c), _p(lr_c)
        if d_u is not None:
            run.uniforms = d_u.data_ptr() + 4 * k * n * uniforms.shape[2]
            run.slots = uniforms.shape[2]
        else:
            run.slots = slots
        run.stream_seed, run.t0, run.agent0, run.env_stream_seed, run.env_t0 = seed, t0 + k, 0, seed, t0 + k
        run.empty_all, run.use_masks = empty_all, use_masks
        if record:
            run.trace_actions = tr["actions"].data_ptr() + 4 * k * n
            run.trace_rewards = tr["rewards"].data_ptr() + 4 * k * n
            run.trace_terminated = tr["terminated"].data_ptr() + k * n
            run.trace_next_states = tr["obs"].data_ptr() + 4 * k * n
            run.trace_episode_returns = tr["episode_returns"].data_ptr() + 4 * k * n
        run.episode_sum, run.episode_count = ep_sum.data_ptr(), ep_cnt.data_ptr()
        capi.check(capi.lib().qe_fused_steps(e.h, C.byref(ag), C.byref(run), None))
        capi.check(capi.lib().qe_sync(e.h, None))
        k += c
    out = {k2: v.cpu().numpy() for k2, v in tr.items()}
    out["ep_sum"], out["ep_count"] = float(ep_sum.item()), int(ep_cnt.item())
    return out


@pytest.mark.parametrize("name", _cases(TTT))
@pytest.mark.parametrize("chunk", [None, 7])
def test_fused_tictactoe_matches_reference_trajectory(capi, name, chunk):
    g = lambda k: TTT[f"{name}__{k}"]  # noqa: E731
    u = g("u_steps")
    steps, n = u.shape[:2]
    dev = torch.device("cuda:0")
    e = Engine(capi, 19683, 9, 0.99)
    try:
        boards = torch.empty(n, dtype=torch.int32, device=dev)
        states = torch.empty(n, dtype=torch.int32, device=dev)
        masks = torch.empty(n, dtype=torch.int32, device=dev)
        d_ui = torch.from_numpy(np.ascontiguousarray(g("u_init")).view(np.int32)).to(dev)
        capi.check(capi.lib().qe_ttt_reset(boards.data_ptr(), states.data_ptr(), masks.data_ptr(), d_ui.data_ptr(), 5, 0, 0, 0, n, None))
        ep_ret = torch.zeros(n, dtype=torch.float32, device=dev)
        decay = float(g("decay"))
        th, lrs = _schedule_arrays(Exponential(0.1, 1e-5, decay), Exponential(1.0, 0.01, decay), steps, n)
        tr = _fused(capi, e, capi.QE_ENV_TTT, n, states, boards, ep_ret, steps=steps, th=th, lrs=lrs, uniforms=u, chunk=chunk)
        for k in ("actions", "rewards", "obs"):
            np.testing.assert_array_equal(tr[k], g(k), err_msg=k)
        np.testing.assert_array_equal(tr["terminated"].astype(bool), g("terminated"))
        er = tr["episode_returns"].ravel()
        np.testing.assert_array_equal(er[~np.isnan(er)], g("history"))
        assert tr["ep_count"] == g("history").shape[0]
        np.testing.assert_array_equal(e.download(), g("q"))  # bit exact vs the real reference (fp32 rewards)
        np.testing.assert_allclose(e.download(), g("q_f64r"), rtol=1e-6, atol=1e-9)  # reference as-is
        np.testing.assert_array_equal(states.cpu().numpy(), g("obs")[-1])
    finally:
        e.close()


@pytest.mark.parametrize("name", _cases(MDP))
def test_fused_mdp_matches_reference_trajectory(capi, name):
    g = lambda k: MDP[f"{name}__{k}"]  # noqa: E731
    s, a, n, steps, seed = (int(x) for x in g("cfg"))
    dev = torch.device("cuda:0")
    e = Engine(capi, s, a, 0.99)
    try:
        states = torch.empty(n, dtype=torch.int32, device=dev)
        d_ui = torch.from_numpy(np.ascontiguousarray(g("u_init")).view(np.int32)).to(dev)
        capi.check(capi.lib().qe_mdp_reset(states.data_ptr(), None, s, a, seed, d_ui.data_ptr(), 4, 0, 0, 0, n, None))
        ep_ret = torch.zeros(n, dtype=torch.float32, device=dev)
        th, lrs = _schedule_arrays(Exponential(0.5, 1e-3, 0.9999), Exponential(1.0, 0.05, 0.9995), steps, n)
        tr = _fused(capi, e, capi.QE_ENV_MDP, n, states, None, ep_ret, steps=steps, th=th, lrs=lrs, uniforms=g("u_steps"),
                    env_seed=seed, term_thresh=int(np.ceil(0.05 * 2.0**32)), empty_all=int(a > 10))
        for k in ("actions", "rewards", "obs"):
            np.testing.assert_array_equal(tr[k], g(k), err_msg=k)
        er = tr["episode_returns"].ravel()
        np.testing.assert_array_equal(er[~np.isnan(er)], g("history"))
        np.testing.assert_array_equal(e.download(), g("q"))
    finally:
        e.close()


@pytest.mark.parametrize("s,a,n,steps", [(1000, 16, 4096, 30), (50_000, 8, 65_536, 12), (1_000_000, 16, 1 << 18, 6), (300, 16, 8192, 300)])
def test_fused_mdp_counter_stream_vs_c_oracle(capi, s, a, n, steps):
    """Heavier collision regimes (N >> S, N ~ S), on-device counter stream, > 255 steps (tag wrap)."""
    seed = 5
    dev = torch.device("cuda:0")
    term_thresh = int(np.ceil(0.05 * 2.0**32))
    th = np.full(steps, orng.explore_threshold(0.1), dtype=np.uint64)
    lrs = np.full(steps, 0.1, dtype=np.float32)
    # oracle
    u_init = orng.draw_uniforms(seed, T_INIT, 1, n, 4)[0]
    st_o, mk_o = co.mdp_reset(u_init, s, a, seed)
    q_o = np.zeros((s, a), dtype=np.float32)
    res = co.run(co.ENV_MDP, q_o, None, st_o, mk_o, num_states=s, env_seed=seed, term_thresh=term_thresh, uniforms=None,
                 stream_seed=seed, steps=steps, eps_thresh=th, lr=lrs, gamma=0.99, empty_all=a > 10)
    assert res["rc"] == 0
    # engine
    e = Engine(capi, s, a, 0.99)
    try:
        states = torch.empty(n, dtype=torch.int32, device=dev)
        capi.check(capi.lib().qe_mdp_reset(states.data_ptr(), None, s, a, seed, None, 4, seed, T_INIT, 0, n, None))
        ep_ret = torch.zeros(n, dtype=torch.float32, device=dev)
        tr = _fused(capi, e, capi.QE_ENV_MDP, n, states, None, ep_ret, steps=steps, th=th, lrs=lrs, uniforms=None, seed=seed,
                    env_seed=seed, term_thresh=term_thresh, empty_all=int(a > 10), record=False)
        np.testing.assert_array_equal(states.cpu().numpy(), st_o)
        np.testing.assert_array_equal(e.download(), q_o)
        np.testing.assert_array_equal(ep_ret.cpu().numpy(), res["agent_rewards"])
        assert tr["ep_count"] == res["ep_count"]
        assert abs(tr["ep_sum"] - res["ep_sum"]) <= 1e-6 * max(1.0, abs(res["ep_sum"]))
    finally:
        e.close()
