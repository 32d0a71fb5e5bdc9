"""Quick device-only timing of the fused loop at the C-ABI level (development aid)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

from dist_classicrl_b200 import capi

S, A, N, K = int(float(sys.argv[1])), int(sys.argv[2]), int(float(sys.argv[3])), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
lib = capi.lib()
h = C.c_void_p()
capi.check(lib.qe_create(S, A, 0.99, 0, C.byref(h)))
capi.check(lib.qe_table_fill_random(h, 1, None))
dev = torch.device("cuda:0")
states = torch.empty(N, dtype=torch.int32, device=dev)
scratch = torch.empty_like(states)
ep = torch.zeros(N, dtype=torch.float32, device=dev)
capi.check(lib.qe_mdp_reset(states.data_ptr(), None, S, A, 0, None, 4, 0, 0xFFFFFFFF, 0, N, None))
th = np.full(K, int(np.ceil(0.1 * 2**32)), dtype=np.uint64)
lr = np.full(K, 0.1, dtype=np.float32)
ag = capi.QeAgents(capi.QE_ENV_MDP, N, states.data_ptr(), scratch.data_ptr(), None, ep.data_ptr(), 0, 0, int(np.ceil(0.05 * 2**32)))
es, ec = torch.zeros(1, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.int64, device=dev)


def launch(t0, stats=True):
    run = capi.QeRun()
    run.steps = K
    run.explore_thresholds_host = th.ctypes.data_as(C.c_void_p)
    run.learning_rates_host = lr.ctypes.data_as(C.c_void_p)
    run.slots = 4
    run.t0 = run.env_t0 = t0
    run.use_masks = 1
    run.evaluate = int(os.environ.get('QE_EVAL', '0'))
    run.learn_mode = int(os.environ.get('QE_ACC', '0'))
    run.empty_all = int(A > 10)
    if stats:
        run.episode_sum, run.episode_count = es.data_ptr(), ec.data_ptr()
    capi.check(lib.qe_fused_steps(h, C.byref(ag), C.byref(run), None))


skip = int(os.environ.get('QE_SKIP', '0'))
for j in range(0, skip, K):
    launch(j)
launch(skip)
capi.check(lib.qe_sync(h, None))
best = 1e9
for r in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    launch(skip + (r + 1) * K)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    best = min(best, ms)
    print(f"rep {r}: {ms:.3f} ms for {K} steps -> {ms / K * 1e3:.1f} us/step, {N * K / ms / 1e6:.3f} G agent-steps/s")
capi.check(lib.qe_sync(h, None))
buf = (C.c_uint64 * 48)()
m = lib.qe_fused_phase_ns(h, buf, 48)
ts = [buf[i] for i in range(m)]
if m >= 4:
    for ph, name in enumerate(("A ", "B1", "B2")):
        print(f"phase {name} us:", " ".join(f"{(ts[1 + ph + 3 * k] - ts[ph + 3 * k]) / 1e3:.1f}" for k in range((m - 1) // 3)))
if lib.qe_fused_form(h) == 3 and m >= 4:
    print("  pipelined form: A = select + env step, B1 = target pipeline, B2 = commit + sort of the next states; commit us:",
          " ".join(f"{(buf[32 + k] - ts[2 + 3 * k]) / 1e3:.1f}" for k in range((m - 1) // 3)), "| sort us:", " ".join(f"{(ts[3 + 3 * k] - buf[32 + k]) / 1e3:.1f}" for k in range((m - 1) // 3)))
if lib.qe_fused_form(h) == 5 and m >= 4:
    print("  one-pass form: A = select + step + targets (+ bucket counts), B1 = column scan + commit, B2 = sort of the next states; scatter us:",
          " ".join(f"{(buf[32 + k] - ts[2 + 3 * k]) / 1e3:.1f}" for k in range((m - 1) // 3)), "| bucket sorts us:", " ".join(f"{(ts[3 + 3 * k] - buf[32 + k]) / 1e3:.1f}" for k in range((m - 1) // 3)))
cnt = (C.c_uint64 * 56)()
if lib.qe_fused_form(h) == 5 and lib.qe_debug_counters(h, cnt, 2) == 0 and cnt[16]:  # only a -DQE_FLOW_STATS build counts
    print('in-order pass (launch totals, warp-passes; lane counts / 32): passes %d busy %d progress %d blocked %d fresh %d produce %d' % tuple(cnt[16 + j] for j in range(6)))
if lib.qe_fused_form(h) == 3:
    if lib.qe_debug_counters(h, cnt, 2) == 0 and cnt[0]:
        print("sort laps of block 0, us per sort (hist, barrier, column scan, barrier, bases, scatter, barrier): pass 0", [round(cnt[8 + j] / 1e3 / K, 1) for j in range(7)],
              "pass 1", [round(cnt[16 + j] / 1e3 / K, 1) for j in range(7)], "segment bounds", round(cnt[24] / 1e3 / K, 1))
        warps = cnt[0]
        print(f"pipeline stats (last launch): warp-steps {warps}, mean time a warp spends in phase T {cnt[1] * 16 / warps / 1e3:.2f} us (slowest {cnt[6] * 16 / 1e3:.1f}), sort {cnt[2] / max(K - 1, 1) / 1e3:.1f} us per step, "
              f"failed polls per agent {cnt[3] / (N * K):.3f}, loop passes per warp-step {cnt[4] / warps:.1f}")
elif lib.qe_debug_counters(h, cnt, 1) == 0 and cnt[5]:
    nw = cnt[5]
    print(f'phase Q per warp-step: sweeps before resident {cnt[0]/nw:.2f}, us until resident mean {cnt[1]/nw/1e3:.1f} max {cnt[2]/1e3:.1f}, resident us mean {cnt[3]/nw/1e3:.1f} max {cnt[4]/1e3:.1f}, jobs at switch {cnt[6]/nw:.1f}, warp-steps {nw}')
balg = 8 * A + 12
print(f"grid={lib.qe_fused_grid_blocks(h)} best {N * K / best / 1e6:.3f} G agent-steps/s, alg {balg} B/agent-step -> {N * K * balg / best / 1e6:.1f} GB/s "
      f"({N * K * balg / best / 1e6 / 6549.4 * 100:.1f}% of measured HBM peak); episodes={int(ec.item())}")
