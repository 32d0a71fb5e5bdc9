"""GPU parity of the batched mixed-radix kernels (qe_radix_encode / qe_radix_decode) and the flatten wrappers:
bit-exact against the live reference's outputs (tests/golden/radix.npz), against the oracle at larger sizes, and the
encode -> decode round trip at 2^22 vectors."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import radix as orx  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "radix.npz"))


@pytest.fixture(scope="module")
def utils():
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    from dist_classicrl_b200 import utils

    return utils


@pytest.mark.parametrize("name", ["ttt", "mixed", "one", "wide"])
def test_batch_kernels_match_reference_outputs(utils, name):
    nvec, radix, vecs, codes = G[f"{name}__nvec"], G[f"{name}__radix"], G[f"{name}__vectors"], G[f"{name}__codes"]
    got = utils.encode_multi_discretes(vecs, radix)
    assert got.dtype == np.int64
    np.testing.assert_array_equal(got, codes)
    np.testing.assert_array_equal(utils.decode_to_multi_discretes(nvec, codes.reshape(-1, 1), radix), G[f"{name}__decoded"])
    np.testing.assert_array_equal(utils.decode_to_multi_discretes(nvec, codes, radix), vecs)
    # device tensors stay on the device
    dv = torch.from_numpy(vecs).cuda()
    dc = utils.encode_multi_discretes(dv, radix)
    assert dc.is_cuda and torch.equal(dc.cpu(), torch.from_numpy(codes.astype(np.int64)))
    assert torch.equal(utils.decode_to_multi_discretes(nvec, dc, radix).cpu(), torch.from_numpy(vecs))


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 100_003])
def test_batch_kernels_match_oracle_on_ragged_sizes(utils, n):
    rng = np.random.default_rng(n)
    nvec = np.array([5, 3, 2, 7, 4, 3], dtype=np.int32)
    radix = utils.compute_radix(nvec)
    vecs = (rng.integers(0, 1 << 20, size=(n, 6)) % nvec).astype(np.int32)
    codes = utils.encode_multi_discretes(vecs, radix)
    np.testing.assert_array_equal(codes, orx.encode(vecs, radix).astype(np.int64).reshape(n))
    np.testing.assert_array_equal(utils.decode_to_multi_discretes(nvec, codes, radix), orx.decode(nvec, codes, radix).reshape(n, 6))


def test_round_trip_at_full_size(utils):
    n = 1 << 22
    nvec = np.array([3] * 9, dtype=np.int32)  # the TicTacToe board of the benchmark
    radix = utils.compute_radix(nvec)
    codes = torch.randint(0, 3**9, (n,), device="cuda", dtype=torch.int64)
    boards = utils.decode_to_multi_discretes(nvec, codes, radix)
    assert boards.shape == (n, 9) and int(boards.min()) >= 0 and int(boards.max()) <= 2
    assert torch.equal(utils.encode_multi_discretes(boards, radix), codes)
    sample = codes[:: n // 64].cpu().numpy()
    np.testing.assert_array_equal(boards[:: n // 64].cpu().numpy(), orx.decode(nvec, sample, radix))


def test_argument_errors(utils):
    with pytest.raises(ValueError):
        utils.encode_multi_discretes(np.zeros((4, 3), np.int32), np.array([4, 2, 1, 1]))
    with pytest.raises(ValueError):
        utils.encode_multi_discretes(np.zeros((4, 40), np.int32), np.ones(40, np.int64))  # more than 32 dims
    with pytest.raises(ZeroDivisionError):
        utils.decode_to_multi_discretes(np.array([2, 0]), np.arange(4), np.array([2, 1]))


class GridVecEnv:
    """Vector twin of the fixture generator's GridEnv: [n, dims] observations, [n, dims] actions."""

    def __init__(self, n, obs_nvec, act_nvec, dict_obs, device):
        from dist_classicrl_b200 import spaces

        self.n, self.obs_nvec, self.act_nvec, self.dict_obs, self.device = n, np.asarray(obs_nvec), np.asarray(act_nvec), dict_obs, device
        sub = spaces.MultiDiscrete(obs_nvec)
        self.observation_space = spaces.Dict({"observation": sub, "action_mask": spaces.MultiDiscrete([2] * 3)}) if dict_obs else sub
        self.action_space = spaces.MultiDiscrete(act_nvec)
        self.t, self.seen = 0, []

    def _obs(self, vec):
        vec = (np.asarray(vec) % self.obs_nvec).astype(np.int32)
        vec = np.repeat(vec[None, :], self.n, axis=0) if vec.ndim == 1 else vec
        if self.device:
            vec = torch.from_numpy(vec).cuda()
        return {"observation": vec, "action_mask": np.ones((self.n, 3), dtype=np.int8)} if self.dict_obs else vec

    def reset(self, seed=None, options=None):
        self.t = 0
        return self._obs(np.arange(len(self.obs_nvec))), {}

    def step(self, action):
        action = action.cpu().numpy() if hasattr(action, "is_cuda") else np.asarray(action)
        self.seen.append(action.copy())
        self.t += 1
        vec = np.stack([np.resize(a, len(self.obs_nvec)) + self.t for a in action])
        return self._obs(vec), np.full(self.n, float(self.t)), np.zeros(self.n, bool), np.zeros(self.n, bool), {}


@pytest.mark.parametrize("name,dict_obs", [("wrap_plain", False), ("wrap_dict", True)])
@pytest.mark.parametrize("device", [False, True])
def test_flatten_wrappers_on_a_vector_env_match_the_reference_wrappers(utils, name, dict_obs, device):
    from dist_classicrl_b200.wrappers import FlattenMultiDiscreteActionsWrapper, FlattenMultiDiscreteObservationsWrapper

    n = 5
    env = GridVecEnv(n, [3, 4, 5, 2], [2, 3, 4], dict_obs, device)
    wrapped = FlattenMultiDiscreteObservationsWrapper(FlattenMultiDiscreteActionsWrapper(env))
    n_obs = wrapped.observation_space.spaces["observation"].n if dict_obs else wrapped.observation_space.n
    assert n_obs == int(G[f"{name}__n_obs"]) and wrapped.action_space.n == int(G[f"{name}__n_act"])
    get = (lambda o: o["observation"]) if dict_obs else (lambda o: o)
    host = lambda x: x.cpu().numpy() if hasattr(x, "is_cuda") else np.asarray(x)  # noqa: E731
    obs, _ = wrapped.reset()
    flat = [host(get(obs))]
    for a in G[f"{name}__actions"]:
        acts = np.full(n, a, dtype=np.int64)
        obs, *_ = wrapped.step(torch.from_numpy(acts).cuda() if device else acts)
        flat.append(host(get(obs)))
    flat = np.stack(flat)  # [T+1, n]: every agent replays the reference's single-env episode
    for i in range(n):
        np.testing.assert_array_equal(flat[:, i], G[f"{name}__flat_obs"])
    np.testing.assert_array_equal(np.stack(env.seen)[:, 0, :], G[f"{name}__inner_actions"])


def test_flatten_wrappers_on_a_single_env_match_the_reference_wrappers(utils):
    """One environment, scalar actions and 1-D observations, exactly how the reference stacks them (TPB:109-123)."""
    from dist_classicrl_b200.wrappers import FlattenMultiDiscreteActionsWrapper, FlattenMultiDiscreteObservationsWrapper

    class One(GridVecEnv):
        def _obs(self, vec):
            vec = (np.asarray(vec) % self.obs_nvec).astype(np.int32)
            return {"observation": vec, "action_mask": np.ones(3, dtype=np.int8)} if self.dict_obs else vec

        def step(self, action):
            self.seen.append(np.asarray(action).copy())
            self.t += 1
            return self._obs(np.resize(np.asarray(action), len(self.obs_nvec)) + self.t), float(self.t), False, False, {}

    for name, dict_obs in (("wrap_plain", False), ("wrap_dict", True)):
        env = One(1, [3, 4, 5, 2], [2, 3, 4], dict_obs, False)
        wrapped = FlattenMultiDiscreteObservationsWrapper(FlattenMultiDiscreteActionsWrapper(env))
        obs, _ = wrapped.reset()
        flat = [obs["observation"] if dict_obs else obs]
        for a in G[f"{name}__actions"]:
            obs, *_ = wrapped.step(int(a))
            flat.append(obs["observation"] if dict_obs else obs)
        np.testing.assert_array_equal(flat, G[f"{name}__flat_obs"])
        np.testing.assert_array_equal(np.stack(env.seen), G[f"{name}__inner_actions"])
