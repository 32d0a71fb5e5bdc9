#!/usr/bin/env python
"""Benchmark of the hot path (masked eps-greedy select -> env step -> sequential TD update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4]

One "step" = one vector step of all agents (N_agents agent-steps).  Metric: agent-steps/s (BASELINE.json).
Workload at --gpus 1: BASELINE config 3 -- hash MDP, 1 000 000 states x 16 actions, 2^20 agents, masked actions,
eps 0.1, lr 0.1, gamma 0.99, uniform[0,1) initial table, on-device counter stream.  With --gpus N > 1 (torchrun)
the state-range-sharded table of config 4 (100 M states x 8 actions, 2^22 agents in total).

Prints ONE JSON line (see the contract in the task description): `value` = device-resident throughput
(CUDA events around the engine's kernel launches, L2 flushed between timed steps), `e2e` = the same metric
through the reference-shaped public API (`SingleThreadQLearning.run_steps`) with the step's pre-drawn uniforms
copied host->device from pinned memory and the step's results copied back, `roofline` (HBM), `cpu_baseline`
(C port of the reference loop on the host cores), `clocks`, `gpu_launches`.
`--impl reference` times the CPU restatement of the reference (oracle/c) on the same workload.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (states, actions, agents, description)
    "c3": (1_000_000, 16, 1 << 20, "hash-MDP 1M states x 16 actions, 2^20 agents, masked (BASELINE config 3)"),
    "c4": (100_000_000, 8, 1 << 22, "hash-MDP 100M states x 8 actions, 2^22 agents, state-range sharded (BASELINE config 4)"),
    "c2": (19_683, 9, 128, "TicTacToe 19683 states x 9 actions, 128 agents, masked (BASELINE config 2)"),
}
EPS, LR, GAMMA, P_TERM, ENV_SEED, STREAM_SEED, TABLE_SEED = 0.1, 0.1, 0.99, 0.05, 0, 0, 1


def alg_bytes(actions: int) -> int:
    """Algorithmic bytes per agent-step: two row reads, one cell write, agent state read+write."""
    return 8 * actions + 12


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.rows: list[list[str]] = []
        self.proc = None
        self.idx = gpu_index

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self) -> None:
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        busy = [x for x in sm if smax and x > 0.5 * smax] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- reference / CPU arm
def cpu_port_rate(workload: str, agents: int, steps: int, warm: int = 1):
    """agent-steps/s of the C restatement of the reference loop (oracle/c) on this host's cores."""
    from oracle import c_oracle as co
    from oracle import rng as orng
    from oracle.envs import T_INIT

    s, a, _n, _ = WORKLOADS[workload]
    thresh = np.full(max(steps, warm), orng.explore_threshold(EPS), dtype=np.uint64)
    lrs = np.full(max(steps, warm), LR, dtype=np.float32)
    if workload == "c2":
        boards, states, masks = co.ttt_reset(orng.draw_uniforms(STREAM_SEED, T_INIT, 1, agents, 5)[0])
        q = np.zeros((s, a), dtype=np.float32)
        kind, env_state, slots, empty_all = co.ENV_TTT, boards, 5, False
    else:
        states, masks = co.mdp_reset(orng.draw_uniforms(STREAM_SEED, T_INIT, 1, agents, 4)[0], s, a, ENV_SEED)
        q = np.random.default_rng(TABLE_SEED).random((s, a), dtype=np.float32)
        kind, env_state, slots, empty_all = co.ENV_MDP, None, 4, a > 10
    kw = dict(num_states=s, env_seed=ENV_SEED, term_thresh=int(math.ceil(P_TERM * 2.0**32)), uniforms=None, slots=slots,
              stream_seed=STREAM_SEED, eps_thresh=thresh, lr=lrs, gamma=GAMMA, empty_all=empty_all)
    co.run(kind, q, env_state, states, masks, t0=0, steps=warm, **kw)
    t = time.perf_counter()
    res = co.run(kind, q, env_state, states, masks, t0=warm, steps=steps, **kw)
    dt = time.perf_counter() - t
    assert res["rc"] == 0
    return agents * steps / dt, dt


def python_port_rate(workload: str, agents: int, steps: int):
    """agent-steps/s of the NumPy/Python restatement (the reference's own per-agent Python loop shape)."""
    from oracle import rng as orng
    from oracle import runtime as ort
    from oracle.envs import T_INIT, HashMDPVec, TicTacToeVec

    s, a, _n, _ = WORKLOADS[workload]
    if workload == "c2":
        env, slots = TicTacToeVec(agents), 5
        q = np.zeros((s, a), dtype=np.float32)
    else:
        env, slots = HashMDPVec(agents, s, a, seed=ENV_SEED, p_term=P_TERM), 4
        q = np.random.default_rng(TABLE_SEED).random((s, a), dtype=np.float32)
    states, _ = env.reset(orng.draw_uniforms(STREAM_SEED, T_INIT, 1, agents, slots)[0])
    u = orng.draw_uniforms(STREAM_SEED, 0, steps, agents, slots)
    t = time.perf_counter()
    ort.run_steps(q, GAMMA, env, u, ort.Constant(LR), ort.Constant(EPS), states=states)
    dt = time.perf_counter() - t
    return agents * steps / dt


def run_reference(args) -> dict:
    """`--impl reference`: the CPU restatement of the reference's path, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        sys.exit(0)
    workload = args.workload or ("c3" if args.gpus == 1 else "c4")
    s, a, n, desc = WORKLOADS[workload]
    # bounded sample: the reference loop is sequential in the agents, so per-agent-step cost does not depend on
    # the batch size; cap the agents so that warm-up + K steps stay within a few minutes
    agents = min(n, 1 << 20)
    steps = max(1, args.steps)
    rate, dt = cpu_port_rate(workload, agents, steps, warm=min(args.warmup, 3) or 1)
    cores = os.cpu_count() or 1
    sample = f"{steps} vector steps x {agents} agents of the same workload (C port of the reference loop, OpenMP select+env, sequential learn)"
    return {
        "impl": "reference", "metric": "agent-steps/s", "value": rate, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{workload}: {desc}", "states": s, "actions": a, "agents": n, "sample_agents": agents,
                   "eps": EPS, "lr": LR, "gamma": GAMMA, "p_term": P_TERM},
        "cpu_baseline": {"value": rate, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# --------------------------------------------------------------------------------------------- our arm, 1 GPU
def run_single_gpu(args) -> dict:
    import torch

    from dist_classicrl_b200 import capi
    from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase
    from dist_classicrl_b200.algorithms.runtime import SingleThreadQLearning
    from dist_classicrl_b200.environments import HashMDPVecEnv, TicTacToeVecEnv
    from dist_classicrl_b200.rng import PredrawnUniforms, draw_uniforms, explore_threshold
    from dist_classicrl_b200.schedules import ConstantSchedule

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    workload = args.workload or "c3"
    s, a, n, desc = WORKLOADS[workload]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = capi.lib()
    K, W = args.steps, max(3, args.warmup)

    def make(seed_offset=0):
        algo = OptimalQLearningBase(s, a, GAMMA, seed=STREAM_SEED + seed_offset)
        if workload == "c2":
            env = TicTacToeVecEnv(n, seed=STREAM_SEED + seed_offset, output="torch")
        else:
            algo.fill_random(TABLE_SEED)
            env = HashMDPVecEnv(n, s, a, env_seed=ENV_SEED, p_term=P_TERM, seed=STREAM_SEED + seed_offset, output="torch")
        env.attach(algo)
        env.reset()
        return algo, env

    # ---------------- device-resident throughput: one launch per vector step, L2 flushed in between
    algo, env = make()
    ep_ret = torch.zeros(n, dtype=torch.float32, device=dev)
    ag = env.agents_struct(ep_ret)
    thresh = np.full(1, explore_threshold(EPS), dtype=np.uint64)
    lrs = np.full(1, LR, dtype=np.float32)
    stats = torch.zeros(2, dtype=torch.float64, device=dev)
    ep_cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # 256 MiB > 126 MB L2

    def launch(t):
        run = capi.QeRun()
        run.steps = 1
        run.explore_thresholds_host = thresh.ctypes.data_as(C.c_void_p)
        run.learning_rates_host = lrs.ctypes.data_as(C.c_void_p)
        run.slots = env.slots
        run.stream_seed = run.env_stream_seed = STREAM_SEED
        run.t0 = run.env_t0 = t
        run.use_masks = 1
        run.empty_all = int(a > 10)
        run.episode_sum, run.episode_count = stats.data_ptr(), ep_cnt.data_ptr()
        capi.check(lib.qe_fused_steps(algo.handle, C.byref(ag), C.byref(run), C.c_void_p(stream.cuda_stream)))

    for t in range(W):
        launch(t)
    capi.check(lib.qe_sync(algo.handle, C.c_void_p(stream.cuda_stream)))
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = lib.qe_kernel_launches(algo.handle)
    events = []
    for k in range(K):
        flush.fill_(k & 0xFF)  # evict the table from L2 between timed steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launch(W + k)
        e1.record(stream)
        events.append((e0, e1))
    torch.cuda.synchronize()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in events]
    capi.check(lib.qe_sync(algo.handle, C.c_void_p(stream.cuda_stream)))
    gpu_launches = int(lib.qe_kernel_launches(algo.handle) - launches0)
    total_ms = sum(step_ms)
    value = n * K / (total_ms * 1e-3)

    # ---------------- steady state: K steps in ONE persistent launch (table stays L2-resident, no flush)
    th_k, lr_k = np.full(K, explore_threshold(EPS), dtype=np.uint64), np.full(K, LR, dtype=np.float32)
    run = capi.QeRun()
    run.steps = K
    run.explore_thresholds_host, run.learning_rates_host = th_k.ctypes.data_as(C.c_void_p), lr_k.ctypes.data_as(C.c_void_p)
    run.slots = env.slots
    run.stream_seed = run.env_stream_seed = STREAM_SEED
    run.t0 = run.env_t0 = W + K
    run.use_masks, run.empty_all = 1, int(a > 10)
    run.episode_sum, run.episode_count = stats.data_ptr(), ep_cnt.data_ptr()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    capi.check(lib.qe_fused_steps(algo.handle, C.byref(ag), C.byref(run), C.c_void_p(stream.cuda_stream)))
    e1.record(stream)
    torch.cuda.synchronize()
    capi.check(lib.qe_sync(algo.handle, C.c_void_p(stream.cuda_stream)))
    steady_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    grid_blocks = int(lib.qe_fused_grid_blocks(algo.handle))
    del algo, env

    # ---------------- e2e through the public API: per step H2D of that step's uniforms (pinned) + D2H of the results
    algo, env = make()
    rt = SingleThreadQLearning(algo, ConstantSchedule(LR), ConstantSchedule(EPS))
    rt.history_mode = "summary"
    Ke = min(K, 20)
    slots = env.slots
    u_host = torch.empty((W + Ke, n, slots), dtype=torch.int32).pin_memory()
    u_host.numpy().view(np.uint32)[:] = draw_uniforms(STREAM_SEED, 0, W + Ke, n, slots)
    pre = PredrawnUniforms(u_host.numpy().view(np.uint32))  # no copy: already contiguous uint32 (pinned)
    algo._rng = env._rng = pre
    sd = {"states": None, "infos": {}, "rewards": np.zeros(n, dtype=np.float32)}
    for _ in range(W):
        _, _, _, sd = rt.run_steps(1, env, sd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(Ke):
        _, _, _, sd = rt.run_steps(1, env, sd)  # returns host copies of the agents' running returns
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = n * slots * 4 + 12
    d2h = n * 4 + 16 + 4
    e2e = {"value": n * Ke / e2e_s, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": Ke, "api": "SingleThreadQLearning.run_steps(1, env, state_dict) with PredrawnUniforms in pinned host memory"}

    # ---------------- roofline + CPU baseline
    peak, peak_src = measured_peak()
    balg = alg_bytes(a)
    achieved = n * balg / (total_ms / K * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(workload)
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": "fused_kernel<MDP,4>" if workload != "c2" else "fused_kernel<TTT,4>",
                "algorithmic_bytes_per_agent_step": balg,
                "steady_state_frac": n * K * balg / (steady_ms * 1e-3) / 1e9 / peak}
    cpu_agents = n if workload == "c2" else 1 << 20
    cpu_steps = 2000 if workload == "c2" else 12
    cpu_rate, cpu_dt = cpu_port_rate(workload, cpu_agents, cpu_steps)
    py_agents, py_steps = (n, 50) if workload == "c2" else (1 << 13, 4)
    py_rate = python_port_rate(workload, py_agents, py_steps)
    cpu_baseline = {"value": cpu_rate, "unit": "agent-steps/s", "cores": os.cpu_count() or 1, "kind": "port",
                    "sample": f"{cpu_steps} vector steps x {cpu_agents} agents, C port of the reference loop ({cpu_dt:.1f} s)",
                    "python_port_value": py_rate, "python_port_cores": 1,
                    "python_port_sample": f"{py_steps} vector steps x {py_agents} agents, NumPy/Python restatement (per-agent Python learn loop like the reference)"}
    return {
        "metric": "agent-steps/s", "value": value, "unit": "agent-steps/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{workload}: {desc}", "states": s, "actions": a, "agents": n, "eps": EPS, "lr": LR, "gamma": GAMMA,
                   "p_term": P_TERM, "table_init": "uniform[0,1)" if workload != "c2" else "zeros", "rng": "on-device counter stream",
                   "timing": "CUDA events around each step's launch; L2 flushed (256 MiB write) between timed steps",
                   "steady_state_value": n * K / (steady_ms * 1e-3), "steady_state_ms_per_step": steady_ms / K,
                   "steady_state_note": "same K steps in ONE persistent launch, table L2-resident, no flush",
                   "grid_blocks": grid_blocks},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
        "episodes": int(ep_cnt.item()),
    }


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, *WORKLOADS])
    args = ap.parse_args()
    if args.impl == "reference":
        out = run_reference(args)
    elif args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from dist_classicrl_b200.distributed import bench_sharded

        out = bench_sharded(args)
        if out is None:
            return
    else:
        out = run_single_gpu(args)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
