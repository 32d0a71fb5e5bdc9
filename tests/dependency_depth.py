"""Dependency depth of one vector step of the sequential TD update on the C3 workload (analysis aid, CPU only; lives under
tests/ because it drives the oracle: `python tests/dependency_depth.py <warm-up steps> [agents]`).  Output of the round-1 run:
profiles/r1_dependency_depth.md."""
import sys, math, ctypes as C
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import c_oracle as co, rng as orng
from oracle.envs import T_INIT

S, A, N = 1_000_000, 16, int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
T = int(sys.argv[1])
states, masks = co.mdp_reset(orng.draw_uniforms(0, T_INIT, 1, N, 4)[0], S, A, 0)
q = np.random.default_rng(1).random((S, A), dtype=np.float32)
kw = dict(num_states=S, env_seed=0, term_thresh=int(math.ceil(0.05 * 2.0**32)), uniforms=None, slots=4, stream_seed=0,
          eps_thresh=np.full(T + 1, orng.explore_threshold(0.1), dtype=np.uint64), lr=np.full(T + 1, 0.1, np.float32), gamma=0.99, empty_all=True)
co.run(co.ENV_MDP, q, None, states, masks, t0=0, steps=T, **kw)
s = states.copy()
res = co.run(co.ENV_MDP, q, None, states, masks, t0=T, steps=1, record=True, **kw)
a = res["trace"]["actions"][0]; s2 = res["trace"]["obs"][0]; term = res["trace"]["terminated"][0]
m2 = np.empty(N, np.uint32)
co.lib().orc_mdp_masks(co._p(np.ascontiguousarray(s2)), C.c_int(A), C.c_uint32(0), C.c_int(N), co._p(m2))
cnt = np.bincount(s, minlength=S)
print(f"step {T}: occupied rows {np.count_nonzero(cnt)}, max cluster {cnt.max()}, agents on rows with >4: {(cnt[s] > 4).mean():.3f}, >16: {(cnt[s] > 16).mean():.3f}, >64: {(cnt[s]>64).mean():.3f}")
# (a) agent-level depth: every dependency = 1 hop; (b) row-sequencer depth: same-row chain free, cross-row read = 1 hop
cellL = {}
rowL = {}
lev = np.zeros(N, np.int32); rlev = np.zeros(N, np.int32)
sl, al, s2l, tl, ml = s.tolist(), a.tolist(), s2.tolist(), term.tolist(), m2.tolist()
for i in range(N):
    si, ai = sl[i], al[i]
    d = cellL.get(si * A + ai, 0)
    rd = rowL.get(si, 0)
    if not tl[i]:
        base = s2l[i] * A
        mk = ml[i]
        for b in range(A):
            if (mk >> b) & 1:
                x = cellL.get(base + b, 0)
                if x > d: d = x
        x = rowL.get(s2l[i], -1)
        if x >= 0 and x + 1 > rd: rd = x + 1
    lev[i] = d + 1
    cellL[si * A + ai] = d + 1
    rlev[i] = rd
    rowL[si] = rd
for name, L in (("agent-level", lev), ("row-sequencer hops", rlev)):
    h = np.bincount(L)
    cum = np.cumsum(h[::-1])[::-1]
    print(name, "max", L.max(), "mean", L.mean().round(2), "agents at level >= k:", {k: int(cum[k]) for k in (1, 2, 3, 5, 8, 12, 16, 24, 32, 48, 64, 96) if k < len(cum)})
