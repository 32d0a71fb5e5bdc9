"""The mixed-radix oracle (oracle/radix.py) against fixtures produced by the REAL reference's utils.py and flatten
wrappers (tests/golden/radix.npz, oracle/make_golden_radix.py), plus the host-side halves of the drop-in
(`compute_radix`, the scalar encode / decode, `DummyVecWrapper`) which need no GPU."""
import os

import numpy as np
import pytest

from oracle import radix as orx

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "radix.npz"))
CASES = ["ttt", "mixed", "one", "wide"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_outputs(name):
    nvec, radix, vecs = G[f"{name}__nvec"], G[f"{name}__radix"], G[f"{name}__vectors"]
    np.testing.assert_array_equal(orx.compute_radix(nvec), radix)
    np.testing.assert_array_equal(orx.encode(vecs, radix), G[f"{name}__codes"])
    np.testing.assert_array_equal([orx.encode(v, radix) for v in vecs[:16]], G[f"{name}__codes_single"])
    np.testing.assert_array_equal(orx.decode(nvec, G[f"{name}__codes"], radix), G[f"{name}__decoded"])
    np.testing.assert_array_equal(G[f"{name}__decoded"], vecs)  # the reference round-trips
    np.testing.assert_array_equal(np.stack([orx.decode(nvec, int(c), radix) for c in G[f"{name}__codes"][:16]]), G[f"{name}__decoded_single"])


@pytest.mark.parametrize("name", CASES)
def test_host_side_functions_match_reference_outputs(name):
    from dist_classicrl_b200 import utils

    nvec, radix, vecs = G[f"{name}__nvec"], G[f"{name}__radix"], G[f"{name}__vectors"]
    got = utils.compute_radix(nvec)
    assert got.dtype == np.int32
    np.testing.assert_array_equal(got, radix)
    for v, c, d in zip(vecs[:16], G[f"{name}__codes_single"], G[f"{name}__decoded_single"]):
        assert utils.encode_multi_discrete(v, radix) == c and isinstance(utils.encode_multi_discrete(v, radix), int)
        np.testing.assert_array_equal(utils.decode_to_multi_discrete(nvec, int(c), radix), d)


class _Env:
    def __init__(self, k):
        self.k, self.t = k, 0
        self.observation_space, self.action_space, self.closed = "obs-space", "act-space", False

    def reset(self, seed=None, options=None):
        self.t = 0
        return np.array([self.k, 0]), {"seed": seed}

    def step(self, action):
        self.t += 1
        return np.array([self.k, self.t]), float(action), self.t >= 2, False, {"t": self.t}

    def close(self):
        self.closed = True

    def render(self):
        return None


def test_dummy_vec_wrapper_stacks_and_forwards():  # dummy_vec_wrapper.py:24-101
    from dist_classicrl_b200.utils import _make_dummy_vec_env
    from dist_classicrl_b200.wrappers import DummyVecWrapper

    env = DummyVecWrapper([_Env(k) for k in range(3)])
    assert env.num_envs == 3 and env.observation_space == "obs-space" and env.k == 0
    obs, infos = env.reset(seed=7)
    np.testing.assert_array_equal(obs, [[0, 0], [1, 0], [2, 0]])
    assert infos == [{"seed": 7}] * 3
    obs, r, term, trunc, infos = env.step(np.array([1, 0, 1]))
    np.testing.assert_array_equal(obs, [[0, 1], [1, 1], [2, 1]])
    np.testing.assert_array_equal(r, [1.0, 0.0, 1.0])
    assert not term.any() and not trunc.any() and infos[2] == {"t": 1}
    _, _, term, _, _ = env.step([0, 0, 0])
    assert term.all()  # no autoreset: the flags come straight from the wrapped environments
    with pytest.raises(ValueError):
        env.step([0, 0])
    env.close()
    assert all(e.closed for e in env.envs)
    made = _make_dummy_vec_env(4, _Env, {"k": 5})
    assert isinstance(made, DummyVecWrapper) and made.num_envs == 4
