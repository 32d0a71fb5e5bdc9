"""Dense-collision probe: fused MDP vs C oracle, prints timing and first mismatch (development aid)."""
import ctypes as C, sys, time
import numpy as np, torch
from dist_classicrl_b200 import capi
from oracle import c_oracle as co, rng as orng
from oracle.envs import T_INIT
S, A, N, K = (int(float(x)) for x in sys.argv[1:5])
seed = 5
lib = capi.lib()
h = C.c_void_p(); capi.check(lib.qe_create(S, A, 0.99, 0, C.byref(h)))
dev = torch.device("cuda:0")
states = torch.empty(N, dtype=torch.int32, device=dev); scratch = torch.empty_like(states)
ep = torch.zeros(N, dtype=torch.float32, device=dev)
capi.check(lib.qe_mdp_reset(states.data_ptr(), None, S, A, seed, None, 4, seed, T_INIT, 0, N, None))
tt = int(np.ceil(0.05 * 2**32))
th = np.full(K, orng.explore_threshold(0.1), dtype=np.uint64); lr = np.full(K, 0.1, dtype=np.float32)
ag = capi.QeAgents(capi.QE_ENV_MDP, N, states.data_ptr(), scratch.data_ptr(), None, ep.data_ptr(), seed, 0, tt)
run = capi.QeRun(); run.steps = K
run.explore_thresholds_host = th.ctypes.data_as(C.c_void_p); run.learning_rates_host = lr.ctypes.data_as(C.c_void_p)
run.slots = 4; run.stream_seed = run.env_stream_seed = seed; run.use_masks = 1; run.empty_all = int(A > 10)
t = time.perf_counter()
rc = lib.qe_fused_steps(h, C.byref(ag), C.byref(run), None)
rc2 = lib.qe_sync(h, None)
dt = time.perf_counter() - t
print("rc", rc, rc2, lib.qe_last_error(), f"{dt*1e3:.1f} ms total, {dt/K*1e6:.1f} us/step")
q = np.empty((S, A), dtype=np.float32); lib.qe_table_download_host(h, q.ctypes.data_as(C.c_void_p))
st_o, mk_o = co.mdp_reset(orng.draw_uniforms(seed, T_INIT, 1, N, 4)[0], S, A, seed)
q_o = np.zeros((S, A), dtype=np.float32)
co.run(co.ENV_MDP, q_o, None, st_o, mk_o, num_states=S, env_seed=seed, term_thresh=tt, uniforms=None, stream_seed=seed,
       steps=K, eps_thresh=th, lr=lr, gamma=0.99, empty_all=A > 10)
print("states equal", np.array_equal(states.cpu().numpy(), st_o), "q equal", np.array_equal(q, q_o), "max|dq|", np.abs(q - q_o).max())
