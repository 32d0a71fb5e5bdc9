#!/usr/bin/env python
"""Benchmark of the hot path (masked eps-greedy select -> env step -> sequential TD update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4]

One "step" = one vector step of all agents (N_agents agent-steps).  Metric: agent-steps/s (BASELINE.json).
Workload at --gpus 1: BASELINE config 3 -- hash MDP, 1 000 000 states x 16 actions, 2^20 agents, masked actions,
eps 0.1, lr 0.1, gamma 0.99, uniform[0,1) initial table, on-device counter stream.  With --gpus N > 1 (torchrun,
one rank per GPU) every GPU runs that workload on its own replica of the table and the replicas are merged by a
Q-delta all-reduce every 8 vector steps (config 5; weak scaling); a bounded run of the state-range-sharded table of
config 4 (100 M states x 8 actions, 2^22 agents in total) is reported beside it under `sharded_c4`.

Prints ONE JSON line (see the contract in the task description): `value` = device-resident throughput
(CUDA events around the K timed steps, 8 vector steps per fused launch; the working set is larger than L2), `e2e` = the same metric
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
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread about every millisecond
    (the timed region of the default run is ~10 ms long, too short for `nvidia-smi -lms`), nvidia-smi as the fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.rows: list[list[str]] = []
        self.proc = None
        self.idx = gpu_index
        self.nvml = None
        self.samples: list[tuple[int, int]] = []  # (sm MHz, reasons bit mask)
        self.smax = None
        self._stop = threading.Event()
        self._thread = None
        self.period = float(os.environ.get("BENCH_CLOCK_PERIOD_MS", "1.0")) * 1e-3  # NVML polling period

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.idx])
            except (ValueError, IndexError):
                pass
        return self.idx

    def prepare(self) -> None:
        """The slow part (NVML initialisation takes milliseconds): call it BEFORE the barrier that opens the timed region."""
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml = (pynvml, h)
            self.sample_now()
            self.samples.clear()
        except Exception:  # noqa: BLE001
            self.nvml = None

    def start(self) -> None:
        if self.nvml is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self) -> None:
        pynvml, h = self.nvml
        reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append((int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(reasons_fn(h))))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def sample_now(self) -> None:
        """One sample from the calling thread."""
        if self.nvml is None:
            return
        pynvml, h = self.nvml
        reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            self.samples.append((int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(reasons_fn(h))))
        except Exception:  # noqa: BLE001
            pass

    def _read(self) -> None:
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.nvml is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            pynvml = self.nvml[0]
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            sm = [float(c) for c, _ in self.samples]
            reasons = sorted({name for _, r in self.samples for name, b in bits.items() if r & b})
            busy = [x for x in sm if self.smax and x > 0.5 * self.smax] or sm
            try:
                pynvml.nvmlShutdown()
            except Exception:  # noqa: BLE001
                pass
            return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": self.smax, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
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
                "samples": len(sm), "source": "nvidia-smi"}


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


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_rate(workload: str, agents: int, steps: int, warm: int = 1):
    """agent-steps/s of the UNMODIFIED reference (pip-installed from /root/reference into baseline/_ref): its own
    ``OptimalQLearningBase`` (``choose_actions`` / ``learn``, stock float64 table and stock RNGs) driven by its own
    ``BaseRuntime.run_single_step`` -- the loop body of ``SingleThreadQLearning.run_steps`` (STR:63-64), which itself
    cannot be imported without gymnasium -- on the NumPy twin of the workload's environment.  Single-threaded, like the
    reference.  Returns (rate, seconds, description) or None if the reference is not installed."""
    if not os.path.isdir(os.path.join(REF_DIR, "dist_classicrl")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        from dist_classicrl.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase as RefQL
        from dist_classicrl.algorithms.runtime.base_runtime import BaseRuntime as RefRuntime
        from dist_classicrl.schedules.constant_schedule import ConstantSchedule as RefConstant
    except Exception:  # noqa: BLE001
        return None
    from oracle import rng as orng
    from oracle.envs import T_INIT, HashMDPVec, TicTacToeVec

    s, a, _n, _ = WORKLOADS[workload]

    class Loop(RefRuntime):  # the ABC's three abstract hooks; everything that runs is the reference's
        def init_training(self):
            return None

        def run_steps(self, steps, env, curr_state_dict):
            return None

        def close_training(self):
            return None

    class Env:  # gym-style step(actions) on top of the oracle environment, one row of the uniform stream per step
        def __init__(self):
            self.inner = TicTacToeVec(agents) if workload == "c2" else HashMDPVec(agents, s, a, seed=ENV_SEED, p_term=P_TERM)
            self.slots = 5 if workload == "c2" else 4
            self.t = 0
            self.num_envs = agents

        def reset(self):
            return self.inner.reset(orng.draw_uniforms(STREAM_SEED, T_INIT, 1, agents, self.slots)[0])

        def step(self, actions):
            u = orng.draw_uniforms(STREAM_SEED, self.t, 1, agents, self.slots)[0]
            self.t += 1
            return self.inner.step(actions, u)

    algo = RefQL(s, a, GAMMA, seed=STREAM_SEED)
    if workload != "c2":
        algo.q_table = np.random.default_rng(TABLE_SEED).random((s, a))
    rt = Loop(algo, RefConstant(LR), RefConstant(EPS))
    env = Env()
    states, _ = env.reset()
    rewards = np.zeros(agents, dtype=np.float32)
    history: list = []
    for _ in range(warm):
        states, _ = rt.run_single_step(env, states, rewards, history)
    t = time.perf_counter()
    for _ in range(steps):
        states, _ = rt.run_single_step(env, states, rewards, history)
    dt = time.perf_counter() - t
    return agents * steps / dt, dt, ("unmodified reference from baseline/_ref: OptimalQLearningBase.choose_actions/learn driven by "
                                    "BaseRuntime.run_single_step, stock float64 table and RNGs, NumPy twin of the environment")


def run_reference(args) -> dict:
    """`--impl reference`: the CPU restatement of the reference's path, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        sys.exit(0)
    workload = args.workload or "c3"
    s, a, n, desc = WORKLOADS[workload]
    # bounded sample: the reference loop is sequential in the agents, so per-agent-step cost does not depend on
    # the batch size; cap the agents so that warm-up + K steps stay within a few minutes
    steps = max(1, args.steps)
    # the reference itself (single-threaded Python: ~25 us per agent-step whatever the batch size) on a bounded sample
    ref_agents = min(n, 1 << 12)
    ref = reference_rate(workload, ref_agents, steps, warm=min(args.warmup, 2) or 1)
    agents = min(n, 1 << 20)
    port_rate, port_dt = cpu_port_rate(workload, agents, min(steps, 12), warm=1)
    cores = os.cpu_count() or 1
    port_sample = f"{min(steps, 12)} vector steps x {agents} agents of the same workload (C port of the reference loop, OpenMP select+env, sequential learn)"
    if ref is not None:
        rate, dt, how = ref
        return {
            "impl": "reference", "metric": "agent-steps/s", "value": rate, "unit": "agent-steps/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{workload}: {desc}", "states": s, "actions": a, "agents_per_gpu": n, "agents": n * args.gpus,
                       "sample_agents": ref_agents, "eps": EPS, "lr": LR, "gamma": GAMMA, "p_term": P_TERM,
                       "ms_per_full_step_scaled": dt / steps * 1e3 * (n / ref_agents),
                       "note": "a step of this arm is one vector step of the SAMPLE (sample_agents agents); the reference's cost per agent-step does not depend on the batch size"},
            "cpu_baseline": {"value": rate, "unit": "agent-steps/s", "cores": 1, "kind": "reference",
                             "sample": f"{steps} vector steps x {ref_agents} agents of the same workload; {how}",
                             "c_port_value": port_rate, "c_port_cores": cores, "c_port_sample": port_sample},
            "e2e": {"value": rate, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
    rate, dt, sample = port_rate, port_dt, port_sample
    steps = min(steps, 12)
    return {
        "impl": "reference", "metric": "agent-steps/s", "value": rate, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{workload}: {desc}", "states": s, "actions": a, "agents_per_gpu": n, "agents": n * args.gpus, "sample_agents": agents,
                   "eps": EPS, "lr": LR, "gamma": GAMMA, "p_term": P_TERM},
        "cpu_baseline": {"value": rate, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# --------------------------------------------------------------------------------------------- our arm
SYNC_EVERY = 8  # vector steps per fused launch; with N > 1 GPUs also the period of the Q-delta all-reduce


def run_ours(args) -> dict | None:
    """N = 1: BASELINE config 3 on one GPU.  N > 1 (torchrun, one rank per GPU): the same workload on EVERY GPU with a
    replicated table merged by a Q-delta all-reduce every SYNC_EVERY steps (config 5, weak scaling), plus a bounded
    run of the state-range-sharded 100M-state table (config 4) reported under `sharded_c4`."""
    import torch

    from dist_classicrl_b200 import capi
    from dist_classicrl_b200 import distributed as D
    from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase
    from dist_classicrl_b200.algorithms.runtime import SingleThreadQLearning
    from dist_classicrl_b200.environments import HashMDPVecEnv, TicTacToeVecEnv
    from dist_classicrl_b200.rng import PredrawnUniforms, draw_uniforms, explore_threshold
    from dist_classicrl_b200.schedules import ConstantSchedule

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    tp = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
        tp = D.TorchDistTransport()
    workload = args.workload or "c3"
    s, a, n, desc = WORKLOADS[workload]
    lib = capi.lib()
    # the engine times its first launches of both forms of the TD update before it settles on one: give it five
    # launches of warm-up at least
    K, W = args.steps, max(3, args.warmup, 5 * SYNC_EVERY)
    # 128 TicTacToe agents take ~11 us per vector step: only long launches amortise the ~40 us a launch costs
    per_launch = 256 if (workload == "c2" and world == 1) else SYNC_EVERY
    stream = torch.cuda.current_stream()

    def sync_all():
        torch.cuda.synchronize()
        if tp is not None:
            tp.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if tp is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if tp is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
        return float(t.item())

    def make():
        algo = OptimalQLearningBase(s, a, GAMMA, seed=STREAM_SEED, device=local)
        if workload == "c2":
            env = TicTacToeVecEnv(n, seed=STREAM_SEED, device=local, output="torch")
        else:
            algo.fill_random(TABLE_SEED)
            env = HashMDPVecEnv(n, s, a, env_seed=ENV_SEED, p_term=P_TERM, seed=STREAM_SEED, device=local, output="torch")
        env.agent0 = rank * n  # every GPU drives its own agents (global agent ids rank*n ...)
        env.attach(algo)
        env.reset()
        return algo, env

    # ---------------- device-resident throughput: SYNC_EVERY vector steps per launch (+ table merge when N > 1)
    algo, env = make()
    if workload == "c2" and "QE_SORTED" not in os.environ:
        capi.check(lib.qe_set_fused_form(algo.handle, 0))  # a per-step sort of 128 agents is all barrier
    ep_ret = torch.zeros(n, dtype=torch.float32, device=dev)
    ag = env.agents_struct(ep_ret)
    stats = torch.zeros(1, dtype=torch.float64, device=dev)
    ep_cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    rt0 = SingleThreadQLearning(algo, ConstantSchedule(LR), ConstantSchedule(EPS))
    rep = D.ReplicatedQLearning(rt0, tp, sync_every=SYNC_EVERY) if tp is not None else None
    t_next = [0]

    learn_mode = [capi.QE_LEARN_SEQUENTIAL]

    def launch(k):
        th = np.full(k, explore_threshold(EPS), dtype=np.uint64)
        lrs = np.full(k, LR, dtype=np.float32)
        run = capi.QeRun()
        run.steps = k
        run.explore_thresholds_host = th.ctypes.data_as(C.c_void_p)
        run.learning_rates_host = lrs.ctypes.data_as(C.c_void_p)
        run.slots = env.slots
        run.stream_seed = run.env_stream_seed = STREAM_SEED
        run.t0 = run.env_t0 = t_next[0]
        run.agent0 = env.agent0
        run.use_masks = 1
        run.empty_all = int(a > 10)
        run.learn_mode = learn_mode[0]
        run.episode_sum, run.episode_count = stats.data_ptr(), ep_cnt.data_ptr()
        capi.check(lib.qe_fused_steps(algo.handle, C.byref(ag), C.byref(run), C.c_void_p(stream.cuda_stream)))
        t_next[0] += k

    def chunks(total):
        out, left = [], total
        while left > 0:
            out.append(min(per_launch, left))
            left -= out[-1]
        return out

    def timed_launch(k) -> float:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        launch(k)
        b.record(stream)
        b.synchronize()
        return a.elapsed_time(b)

    def calibrate(chunk_list) -> dict:
        """The exact TD update has two forms with identical results (writer lists / per-step sort); which one is faster
        depends on how far the agents have herded.  The engine's own selection times its launches as they happen and
        probes the other form every 12 launches -- inside a 5-launch window that is noise -- so the bench decides the
        same way (time each form at the current training progress, keep the faster) but once, during warm-up, and pins
        the result for the window.  Consumes 4 chunks of `chunk_list` (one cold + one timed launch per form)."""
        ms = {}
        for form in (0, 1):
            capi.check(lib.qe_set_fused_form(algo.handle, form))
            launch(chunk_list.pop(0))
            k = chunk_list.pop(0)
            ms[form] = timed_launch(k) / k
        best = 0 if ms[0] <= ms[1] else 1
        capi.check(lib.qe_set_fused_form(algo.handle, best))
        return {"writer_lists_ms_per_step": ms[0], "per_step_sort_ms_per_step": ms[1], "picked": ["writer lists", "per-step sort"][best]}

    calibration = None
    warm = chunks(W)
    if rep is None and workload != "c2" and "QE_SORTED" not in os.environ and len(warm) >= 5:
        calibration = calibrate(warm)
    for k in warm:
        launch(k)
        if rep is not None:
            rep.sync()
    capi.check(lib.qe_sync(algo.handle, C.c_void_p(stream.cuda_stream)))
    if rep is not None and os.environ.get("BENCH_SYNC_TRACE"):
        rep.trace_events = []
    # the sampler starts BEFORE the barrier that opens the timed region: NVML initialisation takes milliseconds, and a
    # rank 0 that enters the region late makes every other rank wait for it at the first all-reduce
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.prepare()
    sync_all()
    if rank == 0:
        sampler.start()
    launches0 = lib.qe_kernel_launches(algo.handle)
    kernel_events = []
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_begin.record(stream)
    # long windows: the workload drifts, so the pinned form is looked at again every 12 launches -- the other form, then
    # the current one, in two adjacent launches of the window (they are ordinary steps of the run), like the engine's
    # own selection does
    cur_form = int(lib.qe_fused_form(algo.handle))
    reprobes = []
    for j, k in enumerate(chunks(K)):
        probing = calibration is not None and j >= 12 and (j % 12) in (0, 1)
        if probing:
            capi.check(lib.qe_set_fused_form(algo.handle, cur_form ^ 1 if j % 12 == 0 else cur_form))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launch(k)
        e1.record(stream)
        kernel_events.append((k, e0, e1))
        if rep is not None:
            rep.sync()
        if probing and j % 12 == 1:
            e1.synchronize()
            (ka, a0, a1), (kb, b0, b1) = kernel_events[-2], kernel_events[-1]
            other, mine = a0.elapsed_time(a1) / ka, b0.elapsed_time(b1) / kb
            if other < mine:
                cur_form ^= 1
            capi.check(lib.qe_set_fused_form(algo.handle, cur_form))
            reprobes.append({"launch": j, "other_ms_per_step": other, "current_ms_per_step": mine, "now": ["writer lists", "per-step sort"][cur_form]})
    e_end.record(stream)
    sync_all()
    capi.check(lib.qe_sync(algo.handle, C.c_void_p(stream.cuda_stream)))
    if rep is not None and getattr(rep, "trace_events", None):
        tr = rep.trace_events
        parts = [sum(e[j].elapsed_time(e[j + 1]) for e in tr) / len(tr) for j in range(3)]
        gaps = [kernel_events[j + 1][1].elapsed_time(kernel_events[j + 1][2]) for j in range(len(kernel_events) - 1)]
        after = [tr[j][3].elapsed_time(kernel_events[j + 1][1]) for j in range(min(len(tr), len(kernel_events) - 1))]
        before = [kernel_events[j][2].elapsed_time(tr[j][0]) for j in range(min(len(tr), len(kernel_events)))]
        sys.stderr.write(f"[rank {rank}] sync: delta {parts[0]:.3f} ms, all-reduce {parts[1]:.3f} ms, merge {parts[2]:.3f} ms; "
                         f"launch->sync gap {sum(before) / len(before):.3f} ms, sync->launch gap {sum(after) / max(1, len(after)):.3f} ms; "
                         f"launches {[round(e0.elapsed_time(e1), 2) for _, e0, e1 in kernel_events]}\n")
    total_ms = max_over_ranks(e_begin.elapsed_time(e_end))
    kernel_ms = sum(e0.elapsed_time(e1) for _, e0, e1 in kernel_events)
    gpu_launches = int(lib.qe_kernel_launches(algo.handle) - launches0)
    value = world * n * K / (total_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    grid_blocks = int(lib.qe_fused_grid_blocks(algo.handle))
    fused_form = ["writer lists", "per-step sort"][int(lib.qe_fused_form(algo.handle))]
    buf = (C.c_uint64 * 33)()
    m = lib.qe_fused_phase_ns(algo.handle, buf, 33)
    phases = None
    if m >= 4:
        ks = (m - 1) // 3
        phases = {name: sum(buf[1 + ph + 3 * j] - buf[ph + 3 * j] for j in range(ks)) / ks / 1e3
                  for ph, name in enumerate(("select_step_register_us", "td_first_pass_us", "td_deferred_us"))}
    # The workload drifts: a greedy policy on a deterministic MDP herds agents onto the same rows, and the rows get more
    # crowded as the table is learned.  Report the same measurement once more after 256 vector steps.
    late = None
    if world == 1 and workload != "c2" and not args.no_late:
        while t_next[0] < 256 - 4 * SYNC_EVERY:
            launch(SYNC_EVERY)
        late_cal = calibrate([SYNC_EVERY] * 4) if calibration is not None else None
        while t_next[0] < 256:
            launch(SYNC_EVERY)
        capi.check(lib.qe_sync(algo.handle, C.c_void_p(stream.cuda_stream)))
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record(stream)
        for _ in range(4):
            launch(SYNC_EVERY)
        l1.record(stream)
        torch.cuda.synchronize()
        capi.check(lib.qe_sync(algo.handle, C.c_void_p(stream.cuda_stream)))
        late_ms = l0.elapsed_time(l1) / (4 * SYNC_EVERY)
        late = {"td_update_form": ["writer lists", "per-step sort"][int(lib.qe_fused_form(algo.handle))], "after_vector_steps": 256, "steps": 4 * SYNC_EVERY, "ms_per_step": late_ms, "value": n / (late_ms * 1e-3), "unit": "agent-steps/s",
                "form_calibration": late_cal}
    episodes = int(sum_over_ranks(float(ep_cnt.item())))
    del algo, env, rep, rt0

    # ---------------- the same loop with the plain-atomics update (learn_vec, QLO:819-891) instead of the exact sequential
    # one: what the ordering guarantee costs on this workload (BASELINE north_star, item 3).  Not the headline: the
    # reference's trainers call learn(), which is sequential.
    atomics = None
    if world == 1 and not args.no_late:
        algo, env = make()
        ep_ret = torch.zeros(n, dtype=torch.float32, device=dev)
        ag = env.agents_struct(ep_ret)
        stats.zero_()
        ep_cnt.zero_()
        t_next[0] = 0
        learn_mode[0] = capi.QE_LEARN_ACCUMULATE
        for k in chunks(W):
            launch(k)
        capi.check(lib.qe_sync(algo.handle, C.c_void_p(stream.cuda_stream)))
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for k in chunks(K):
            launch(k)
        a1.record(stream)
        torch.cuda.synchronize()
        capi.check(lib.qe_sync(algo.handle, C.c_void_p(stream.cuda_stream)))
        learn_mode[0] = capi.QE_LEARN_SEQUENTIAL
        acc_ms = a0.elapsed_time(a1) / K
        m = lib.qe_fused_phase_ns(algo.handle, buf, 33)
        acc_phases = None
        if m >= 4:
            ks = (m - 1) // 3
            acc_phases = {name: sum(buf[1 + ph + 3 * j] - buf[ph + 3 * j] for j in range(ks)) / ks / 1e3
                          for ph, name in enumerate(("select_step_us", "bootstrap_delta_us", "atomic_scatter_us"))}
        peak_gbs = measured_peak()[0]
        atomics = {"td_update": "learn_vec semantics: snapshot bootstrap + atomicAdd scatter (not what the reference's trainers call)",
                   "value": n / (acc_ms * 1e-3), "unit": "agent-steps/s", "steps": K, "ms_per_step": acc_ms, "phase_us_per_step": acc_phases,
                   "roofline_frac": n * alg_bytes(a) / (acc_ms * 1e-3) / 1e9 / peak_gbs, "kernel": "fused_kernel<MDP,2,ACC>" if workload != "c2" else "fused_kernel<TTT,2,ACC>"}
        del algo, env

    # ---------------- e2e through the public API: per step H2D of that step's uniforms (pinned) + D2H of the results
    algo, env = make()
    rt = SingleThreadQLearning(algo, ConstantSchedule(LR), ConstantSchedule(EPS))
    rt.history_mode = "summary"
    runner = D.ReplicatedQLearning(rt, tp, sync_every=SYNC_EVERY, carry_over=True) if tp is not None else rt
    Ke = min(K, 20)
    slots = env.slots
    u_host = torch.empty((W + Ke, n, slots), dtype=torch.int32).pin_memory()
    u_host.numpy().view(np.uint32)[:] = draw_uniforms(STREAM_SEED, 0, W + Ke, n, slots, agent0=env.agent0)
    pre = PredrawnUniforms(u_host.numpy().view(np.uint32))  # no copy: already contiguous uint32 (pinned)
    algo._rng = env._rng = pre
    sd = {"states": None, "infos": {}, "rewards": np.zeros(n, dtype=np.float32)}
    for _ in range(W):
        _, _, _, sd = runner.run_steps(1, env, sd)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(Ke):
        _, _, _, sd = runner.run_steps(1, env, sd)  # returns host copies of the agents' running returns
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = n * slots * 4 + n * 4 + 12  # the step's uniforms + the agents' running returns (state dict) + eps/lr of the step
    d2h = n * 4 + 16                  # the running returns + {sum, count} of the episodes that finished in the step
    e2e = {"value": world * n * Ke / e2e_s, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
           "steps": Ke, "api": (f"ReplicatedQLearning(sync_every={SYNC_EVERY}, carry_over=True)." if tp is not None else "SingleThreadQLearning.") +
           "run_steps(1, env, state_dict): the step's pre-drawn uniforms (PredrawnUniforms, pinned host memory) and the state dict's running returns go host->device, the returns and the episode statistics come back, every step"}
    del algo, env, runner, rt

    # ---------------- sharded 100M-state table (config 4), bounded
    sharded = None
    if tp is not None and not args.no_sharded:
        s4, a4, n4, desc4 = WORKLOADS["c4"]
        sh = D.ShardedQLearning(s4, a4, GAMMA, n4, tp, env_seed=ENV_SEED, p_term=P_TERM, seed=STREAM_SEED, device=local)
        sh.fill_random(TABLE_SEED)
        sh.reset()
        eps_s, lr_s = ConstantSchedule(EPS), ConstantSchedule(LR)
        sh.run_steps(3, eps_s, lr_s)
        sync_all()
        r0 = sh.rounds_total
        ks = min(K, 10)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        sh.run_steps(ks, eps_s, lr_s)
        b1.record(stream)
        sync_all()
        ms = max_over_ranks(b0.elapsed_time(b1))
        sharded = {"workload": f"c4: {desc4}", "value": n4 * ks / (ms * 1e-3), "unit": "agent-steps/s", "scaling": "strong", "steps": ks,
                   "ms_per_step": ms / ks, "fixed_point_rounds_per_step": (sh.rounds_total - r0) / ks,
                   "states": s4, "actions": a4, "agents": n4,
                   "exchange": "NCCL all-to-all: bootstrap requests + answers per round, agent migration per step"}
        del sh

    if rank != 0:
        return None
    # ---------------- roofline + CPU baseline (rank 0)
    peak, peak_src = measured_peak()
    balg = alg_bytes(a)
    per_launch_ms = kernel_ms / len(kernel_events)
    steps_per_launch = K / len(kernel_events)
    achieved = n * steps_per_launch * balg / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(workload + ("_sorted" if fused_form == "per-step sort" else ""))
        except Exception:  # noqa: BLE001
            traffic = None
    kname = ("fused_sorted_kernel" if fused_form == "per-step sort" else "fused_kernel") + ("<MDP,2>" if workload != "c2" else "<TTT,2>")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": kname, "algorithmic_bytes_per_agent_step": balg,
                "algorithmic_bytes_per_launch": n * steps_per_launch * balg, "avg_launch_ms": per_launch_ms,
                "steps_per_launch": steps_per_launch, "phase_us_per_step": phases}
    cpu_baseline = None
    if world == 1:
        cpu_agents = n if workload == "c2" else 1 << 20
        cpu_steps = 2000 if workload == "c2" else 12
        cpu_rate, cpu_dt = cpu_port_rate(workload, cpu_agents, cpu_steps)
        py_agents, py_steps = (n, 50) if workload == "c2" else (1 << 13, 4)
        py_rate = python_port_rate(workload, py_agents, py_steps)
        ref = reference_rate(workload, min(n, 1 << 12), 4)
        if ref is not None:
            cpu_baseline = {"value": ref[0], "unit": "agent-steps/s", "cores": 1, "kind": "reference",
                            "sample": f"4 vector steps x {min(n, 1 << 12)} agents ({ref[1]:.1f} s); {ref[2]}",
                            "c_port_value": cpu_rate, "c_port_cores": os.cpu_count() or 1,
                            "c_port_sample": f"{cpu_steps} vector steps x {cpu_agents} agents, C port of the reference loop ({cpu_dt:.1f} s)",
                            "python_port_value": py_rate, "python_port_cores": 1,
                            "python_port_sample": f"{py_steps} vector steps x {py_agents} agents, NumPy/Python restatement"}
        else:
            cpu_baseline = {"value": cpu_rate, "unit": "agent-steps/s", "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"{cpu_steps} vector steps x {cpu_agents} agents, C port of the reference loop ({cpu_dt:.1f} s)",
                        "python_port_value": py_rate, "python_port_cores": 1,
                        "python_port_sample": f"{py_steps} vector steps x {py_agents} agents, NumPy/Python restatement (per-agent Python learn loop like the reference)"}
    cfg = {"workload": f"{workload}: {desc}" + (f", one replica per GPU, Q-delta all-reduce every {SYNC_EVERY} steps (BASELINE config 5)" if world > 1 else ""),
           "states": s, "actions": a, "agents_per_gpu": n, "agents": n * world, "eps": EPS, "lr": LR, "gamma": GAMMA, "p_term": P_TERM,
           "table_init": "uniform[0,1)" if workload != "c2" else "zeros", "rng": "on-device counter stream",
           "steps_per_launch": per_launch,
           "timing": "CUDA events around the K timed steps (max over ranks); no L2 flush: the working set (256 MB of row blocks + "
                     "~60 MB of per-agent arrays) is larger than the 126 MB L2",
           "grid_blocks": grid_blocks, "td_update_form_at_end_of_window": fused_form, "timed_window": f"vector steps {W}..{W + K} of the run",
           "td_update_form_calibration": calibration, "td_update_form_reprobes": reprobes or None, "late_training": late}
    out = {
        "metric": "agent-steps/s", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg, "roofline": roofline, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
        "episodes": episodes,
    }
    if atomics is not None:
        out["atomics_mode"] = atomics
    if cpu_baseline is not None:
        out["cpu_baseline"] = cpu_baseline
    if sharded is not None:
        out["sharded_c4"] = sharded
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, *WORKLOADS])
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the bounded run of the sharded 100M-state table")
    ap.add_argument("--no-late", action="store_true", help="N = 1: skip the extra measurement after 256 vector steps")
    args = ap.parse_args()
    if args.impl == "reference":
        out = run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            raise SystemExit(f"--gpus {args.gpus} needs one rank per GPU: launch with python -m torch.distributed.run "
                             f"--nnodes=1 --nproc-per-node {args.gpus} --master-addr 127.0.0.1 bench.py --gpus {args.gpus} ...")
        out = run_ours(args)
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
            dist.destroy_process_group()
        if out is None:
            return
    print(json.dumps(out))


if __name__ == "__main__":
    main()
