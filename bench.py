#!/usr/bin/env python
"""Benchmark of the hot path (masked eps-greedy select -> env step -> sequential TD update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c1|c2|c4]

One "step" = one vector step of all agents (N_agents agent-steps).  Metric: agent-steps/s (BASELINE.json).
Workload at --gpus 1: BASELINE config 3 -- hash MDP, 1 000 000 states x 16 actions, 2^20 agents, masked actions,
eps 0.1, lr 0.1, gamma 0.99, uniform[0,1) initial table, on-device counter stream (`--workload c2` / `c4`: configs 2 and
4 on one GPU).  With --gpus N > 1 (torchrun, one rank per GPU) every GPU runs that workload on its own replica of the
table and the replicas are merged by a Q-delta all-reduce every 8 vector steps (config 5; weak scaling); the
state-range-sharded table of config 4 (100 M states x 8 actions, 2^22 agents in total, peer memory over NVLink) is
reported beside it under `sharded_c4`, and `multi_gpu_parity` says whether the sharded and the replicated mode reproduced
the C oracle on a small case in this very run.

Prints ONE JSON line (see the contract in the task description): `value` = device-resident throughput (CUDA events around
the K timed steps that follow W warm-up steps, 8 vector steps per fused launch), `value_long` = the same over the first
512 vector steps of a fresh run (the workload drifts: agents herd as the table is learned), `e2e` = the same metric
through the reference-shaped public API (`SingleThreadQLearning.run_steps`) with the step's pre-drawn uniforms copied
host->device from pinned memory and the step's results copied back, `roofline` (HBM; `gather_peak` = what dependency-free
random row gathers reach on this table), `cpu_baseline` (the unmodified reference from baseline/_ref: single-thread and
multiprocessing trainers, MPI recorded as not runnable; C and NumPy ports beside it), `clocks`, `gpu_launches`.
`--impl reference` times the unmodified reference (baseline/_ref) on the host cores, on a bounded sample of the workload.
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
    "c1": (19_683, 9, 1, "TicTacToe 19683 states x 9 actions, 1 agent, masked (BASELINE config 1, the reference's own CPU-runnable case)"),
}
SMALL = ("c1", "c2")  # one-CTA kernel, TicTacToe
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
    if workload in SMALL:
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
    if workload in SMALL:
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
FORM_NAMES = {0: "writer lists", 1: "per-step sort", 3: "target pipeline", 4: "one-CTA loop (small batches)", 5: "one-pass pipeline"}
FORM_KERNELS = {0: "fused_kernel", 1: "fused_sorted_kernel", 3: "fused_pipe_kernel", 4: "fused_small_kernel", 5: "fused_flow_kernel"}


class TwinEnv:
    """gym-style ``reset()`` / ``step(actions)`` on top of the oracle's NumPy environment (the reference's trainers need
    gymnasium for their own environments, which is not installed): one row of the uniform stream per step."""

    def __init__(self, workload: str, agents: int, agent0: int = 0):
        from oracle.envs import HashMDPVec, TicTacToeVec

        s, a, _n, _ = WORKLOADS[workload]
        self.inner = TicTacToeVec(agents) if workload in SMALL else HashMDPVec(agents, s, a, seed=ENV_SEED, p_term=P_TERM)
        self.slots = 5 if workload in SMALL else 4
        self.t = 0
        self.num_envs = agents
        self.agent0 = agent0

    def reset(self, seed=None):
        from oracle import rng as orng
        from oracle.envs import T_INIT

        return self.inner.reset(orng.draw_uniforms(STREAM_SEED, T_INIT, 1, self.num_envs, self.slots, agent0=self.agent0)[0])

    def step(self, actions):
        from oracle import rng as orng

        u = orng.draw_uniforms(STREAM_SEED, self.t, 1, self.num_envs, self.slots, agent0=self.agent0)[0]
        self.t += 1
        return self.inner.step(actions, u)


def _import_reference():
    if not os.path.isdir(os.path.join(REF_DIR, "dist_classicrl")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        from dist_classicrl.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase as RefQL
        from dist_classicrl.algorithms.runtime.base_runtime import BaseRuntime as RefRuntime
        from dist_classicrl.algorithms.runtime.parallel_runtime import ParallelQLearning as RefParallel
        from dist_classicrl.schedules.constant_schedule import ConstantSchedule as RefConstant
    except Exception:  # noqa: BLE001
        return None
    return RefQL, RefRuntime, RefParallel, RefConstant


def reference_parallel_rates(workload: str, agents_per_env: int, steps_per_proc: int, procs: list[int]) -> dict:
    """agent-steps/s of the UNMODIFIED reference's multiprocessing trainer (``ParallelQLearning.run_steps``, PRT:80-165:
    one process per environment, the table in shared memory, ONE lock around ``choose_actions`` and ``learn``) with P
    environments of ``agents_per_env`` agents each, the reference's own throughput formula (agent-steps of all processes /
    wall time around ``run_steps``, process start-up included, TPB:222-249)."""
    ref = _import_reference()
    if ref is None:
        return {}
    RefQL, _RefRuntime, RefParallel, RefConstant = ref
    s, a, _n, _ = WORKLOADS[workload]
    out = {}
    for p in procs:
        algo = RefQL(s, a, GAMMA, seed=STREAM_SEED)
        if workload not in SMALL:
            algo.q_table = np.random.default_rng(TABLE_SEED).random((s, a))
        rt = RefParallel(algo, RefConstant(LR), RefConstant(EPS))
        envs = [TwinEnv(workload, agents_per_env, agent0=k * agents_per_env) for k in range(p)]
        try:
            rt.init_training()
            t = time.perf_counter()
            rt.run_steps(steps_per_proc * p, envs, None)
            dt = time.perf_counter() - t
            out[str(p)] = {"value": p * steps_per_proc * agents_per_env / dt, "seconds": dt}
        except Exception as exc:  # noqa: BLE001
            out[str(p)] = {"error": repr(exc)}
        finally:
            try:
                rt.close_training()
            except Exception:  # noqa: BLE001
                pass
    return out


def cpu_baselines(workload: str, ref_steps: int = 4, quick: bool = False) -> dict:
    """Everything the GPU numbers are printed next to, timed on THIS host (no CUDA in this process): the unmodified
    reference single-threaded and through its multiprocessing trainer, its MPI trainer recorded as not runnable, and the
    C / NumPy ports of the same loop."""
    s, a, n, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    out = {"cores_on_host": cores}
    ref_agents = min(n, 1 << 12)
    ref = reference_rate(workload, ref_agents, ref_steps)
    if ref is not None:
        out["single_thread"] = {"value": ref[0], "cores": 1, "sample": f"{ref_steps} vector steps x {ref_agents} agents ({ref[1]:.1f} s); {ref[2]}"}
        procs = sorted({p for p in (1, 2, 4, 8, cores) if p <= cores})
        per_env = min(n, 1 << 11)
        spp = (50 if quick else 200) if workload in SMALL else (2 if quick else 8)
        par = reference_parallel_rates(workload, per_env, spp, procs)
        out["multiprocessing"] = {"trainer": "unmodified reference ParallelQLearning.run_steps (one process per environment, shared-memory table, one lock around choose_actions and learn)",
                                  "agents_per_env": per_env, "vector_steps_per_process": spp, "by_processes": par,
                                  "formula": "agent-steps of all processes / wall time around run_steps (process start-up included, as the reference's throughput_benchmark.py does)"}
        good = [(v["value"], int(p)) for p, v in par.items() if "value" in v]
        if good:
            best = max(good)
            out["multiprocessing"]["best"] = {"value": best[0], "processes": best[1]}
    out["mpi"] = {"value": None, "status": "not runnable here: neither mpi4py nor an MPI launcher is installed (image and wheelhouse)",
                  "published_i7_11700K": {"8_ranks_128_agents": 76098, "2_ranks_128_agents": 22596, "source": "BASELINE.md (benchmark_results/distributed_128_agents_*_processes.json)"}}
    cpu_agents = n if workload in SMALL else 1 << 20
    cpu_steps = 2000 if workload in SMALL else (4 if quick else 12)
    rate, dt = cpu_port_rate(workload if workload != "c4" else "c3", cpu_agents, cpu_steps)
    out["c_port"] = {"value": rate, "cores": cores, "sample": f"{cpu_steps} vector steps x {cpu_agents} agents, C port of the reference loop: OpenMP select + env step, sequential learn ({dt:.1f} s)"
                     + (" [config 3's table: config 4's 100M x 8 fp32 table is 3.2 GB per copy]" if workload == "c4" else "")}
    py_agents, py_steps = (n, 50) if workload in SMALL else (1 << 13, 2 if quick else 4)
    out["python_port"] = {"value": python_port_rate(workload if workload != "c4" else "c3", py_agents, py_steps), "cores": 1,
                          "sample": f"{py_steps} vector steps x {py_agents} agents, NumPy/Python restatement"}
    return out


def reference_rate(workload: str, agents: int, steps: int, warm: int = 1):
    """agent-steps/s of the UNMODIFIED reference (pip-installed from /root/reference into baseline/_ref): its own
    ``OptimalQLearningBase`` (``choose_actions`` / ``learn``, stock float64 table and stock RNGs) driven by its own
    ``BaseRuntime.run_single_step`` -- the loop body of ``SingleThreadQLearning.run_steps`` (STR:63-64), which itself
    cannot be imported without gymnasium -- on the NumPy twin of the workload's environment.  Single-threaded, like the
    reference.  Returns (rate, seconds, description) or None if the reference is not installed."""
    ref = _import_reference()
    if ref is None:
        return None
    RefQL, RefRuntime, _RefParallel, RefConstant = ref
    s, a, _n, _ = WORKLOADS[workload]

    class Loop(RefRuntime):  # the ABC's three abstract hooks; everything that runs is the reference's
        def init_training(self):
            return None

        def run_steps(self, steps, env, curr_state_dict):
            return None

        def close_training(self):
            return None

    algo = RefQL(s, a, GAMMA, seed=STREAM_SEED)
    if workload not in SMALL:
        algo.q_table = np.random.default_rng(TABLE_SEED).random((s, a))
    rt = Loop(algo, RefConstant(LR), RefConstant(EPS))
    env = TwinEnv(workload, agents)
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
    """`--impl reference`: the unmodified reference (baseline/_ref) on this host's cores, bounded sample of the workload:
    its single-thread trainer loop and its multiprocessing trainer with 1, 2, 4, 8, ... processes; the line's value is the
    best of them (the reference "with all the host threads it can use")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        sys.exit(0)
    workload = args.workload or "c3"
    s, a, n, desc = WORKLOADS[workload]
    steps = max(1, min(args.steps, 8))
    base = cpu_baselines(workload, ref_steps=steps, quick=False)
    cores = os.cpu_count() or 1
    cands = []
    if "single_thread" in base:
        cands.append((base["single_thread"]["value"], 1, "single-thread trainer loop: " + base["single_thread"]["sample"]))
    if "multiprocessing" in base and "best" in base["multiprocessing"]:
        mpb = base["multiprocessing"]
        cands.append((mpb["best"]["value"], mpb["best"]["processes"],
                      f"multiprocessing trainer, {mpb['best']['processes']} processes x {mpb['agents_per_env']} agents x {mpb['vector_steps_per_process']} vector steps each; {mpb['trainer']}"))
    kind = "reference"
    if not cands:  # baseline/_ref is missing: the C port stands in
        cands.append((base["c_port"]["value"], cores, base["c_port"]["sample"]))
        kind = "port"
    rate, used, sample = max(cands)
    ref_agents = min(n, 1 << 12)
    return {
        "impl": "reference", "metric": "agent-steps/s", "value": rate, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ref_agents / rate * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if kind == "reference" else "f32", "data": "synthetic",
        "config": {"workload": f"{workload}: {desc}", "states": s, "actions": a, "agents_per_gpu": n, "agents": n * args.gpus,
                   "sample_agents": ref_agents, "eps": EPS, "lr": LR, "gamma": GAMMA, "p_term": P_TERM,
                   "note": "a step of this arm is one vector step of the SAMPLE (sample_agents agents); the reference's cost per agent-step "
                           "does not depend on the batch size (a per-agent Python loop), so agent-steps/s carries over to the full batch"},
        "cpu_baseline": {"value": rate, "unit": "agent-steps/s", "cores": used, "kind": kind, "sample": sample, **{k: v for k, v in base.items()}},
        "e2e": {"value": rate, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def cpu_baselines_subprocess(workload: str) -> dict | None:
    """The CPU arms in a FRESH interpreter (the reference's multiprocessing trainer forks: not from a process that holds a
    CUDA context)."""
    try:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "_cpu", "--workload", workload], capture_output=True, text=True,
                             timeout=900, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
        for line in reversed(res.stdout.splitlines()):
            if line.startswith("{"):
                return json.loads(line)
    except Exception:  # noqa: BLE001
        return None
    return None


def multi_gpu_parity(tp, local: int, torch) -> dict:
    """Sharded and replicated modes on small cases, this very set of GPUs, against the C oracle (rank 0 compares):
    sharded -- table, agent states and running returns bit for bit; replicated -- states bit for bit and the merged table
    against the NumPy restatement of the delta rule (exact for two ranks, 1e-6 relative beyond: the order in which the
    all-reduce adds more than two fp32 terms is NCCL's)."""
    from dist_classicrl_b200 import distributed as D
    from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase
    from dist_classicrl_b200.algorithms.runtime import SingleThreadQLearning
    from dist_classicrl_b200.environments import HashMDPVecEnv
    from dist_classicrl_b200.schedules import ConstantSchedule
    from oracle import c_oracle as co
    from oracle import rng as orng
    from oracle.envs import T_INIT

    tt = int(math.ceil(P_TERM * 2.0**32))
    out = {}

    def rnd_table(S, A, seed):
        x = (np.arange(S * A, dtype=np.uint64) ^ np.uint64((seed * 0x9E3779B9) & 0xFFFFFFFF)) & np.uint64(0xFFFFFFFF)
        x ^= x >> np.uint64(16)
        x = (x * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
        x ^= x >> np.uint64(13)
        x = (x * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
        x ^= x >> np.uint64(16)
        return ((x >> np.uint64(8)).astype(np.float32) * np.float32(2.0**-24)).reshape(S, A)

    # sharded
    S, A, N, steps, seed, env_seed = 300_000, 16, 100_000, 6, 7, 3
    sh = D.ShardedQLearning(S, A, GAMMA, N, tp, env_seed=env_seed, p_term=P_TERM, seed=seed, device=local)
    sh.fill_random(TABLE_SEED)
    sh.reset()
    sh.run_steps(steps, ConstantSchedule(EPS), ConstantSchedule(LR))
    table = sh.gather_table()
    states, rets = sh.gather_agents()
    sh.close()
    if tp.rank == 0:
        st_o, mk_o = co.mdp_reset(orng.draw_uniforms(seed, T_INIT, 1, N, 4)[0], S, A, env_seed)
        q_o = rnd_table(S, A, TABLE_SEED)
        rew = np.zeros(N, dtype=np.float32)
        res = co.run(co.ENV_MDP, q_o, None, st_o, mk_o, num_states=S, env_seed=env_seed, term_thresh=tt, uniforms=None, slots=4, stream_seed=seed,
                     steps=steps, eps_thresh=np.full(steps, orng.explore_threshold(EPS), dtype=np.uint64), lr=np.full(steps, LR, np.float32), gamma=GAMMA,
                     empty_all=A > 10, agent_rewards=rew)
        ok = res["rc"] == 0 and np.array_equal(states, st_o) and np.array_equal(table, q_o) and np.array_equal(rets, rew)
        out["sharded"] = {"result": "pass" if ok else "FAIL", "case": f"{S} states x {A} actions, {N} agents, {steps} vector steps, {tp.world_size} GPUs: table, states, returns bit-exact vs the C oracle"}
    # replicated: one merge period
    S, A, n_local, k, seed, env_seed = 3000, 8, 2048, 4, 11, 2
    algo = OptimalQLearningBase(S, A, GAMMA, seed=seed, device=local)
    algo.fill_random(TABLE_SEED)
    env = HashMDPVecEnv(n_local, S, A, env_seed=env_seed, p_term=P_TERM, seed=seed, device=local, output="torch")
    env.agent0 = tp.rank * n_local
    env.attach(algo)
    rt = SingleThreadQLearning(algo, ConstantSchedule(LR), ConstantSchedule(EPS))
    rt.history_mode = "summary"
    rep = D.ReplicatedQLearning(rt, tp, sync_every=k)
    rep.run_steps(k, env)
    mine = torch.stack([env.states.to(torch.int32)]).reshape(1, -1).contiguous()
    all_states = tp.all_gather_rows(mine).cpu().numpy()
    merged = np.array(algo.q_table, copy=True)
    if tp.rank == 0:
        base = rnd_table(S, A, TABLE_SEED)
        acc = np.zeros_like(base, dtype=np.float64)
        good = True
        for r in range(tp.world_size):
            st_r = co.mdp_reset(orng.draw_uniforms(seed, T_INIT, 1, n_local, 4, agent0=r * n_local)[0], S, A, env_seed)
            q = base.copy()
            res = co.run(co.ENV_MDP, q, None, st_r[0], st_r[1], num_states=S, env_seed=env_seed, term_thresh=tt, uniforms=None, slots=4, stream_seed=seed,
                         t0=0, agent0=r * n_local, steps=k, eps_thresh=np.full(k, orng.explore_threshold(EPS), dtype=np.uint64), lr=np.full(k, LR, np.float32),
                         gamma=GAMMA, empty_all=A > 10)
            good = good and res["rc"] == 0 and np.array_equal(all_states[r], st_r[0])
            acc += (q - base).astype(np.float64)
        want = base.astype(np.float64) + acc
        err = float(np.max(np.abs(merged.astype(np.float64) - want) / np.maximum(1.0, np.abs(want))))
        good = good and err <= 1e-6
        out["replicated"] = {"result": "pass" if good else "FAIL", "max_rel_err_table": err,
                             "case": f"{S} states x {A} actions, {n_local} agents per GPU, one merge after {k} steps, {tp.world_size} GPUs: states bit-exact, merged table within 1e-6 of base + sum of the oracle's deltas"}
        out["result"] = "pass" if all(v.get("result") == "pass" for v in out.values() if isinstance(v, dict)) else "FAIL"
    del algo, env, rep, rt
    return out


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
    K, W = args.steps, max(3, args.warmup)  # (at least three warm-up steps; the first launch also pays for module loading)
    # 128 TicTacToe agents take ~11 us per vector step: only long launches amortise the ~40 us a launch costs
    per_launch = 256 if (workload in SMALL and world == 1) else SYNC_EVERY
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
        if workload in SMALL:
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

    calibration = None  # (round 1 timed two forms of the exact update against each other here; the target pipeline replaced both)
    warm = chunks(W)
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
    reprobes = []
    for j, k in enumerate(chunks(K)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launch(k)
        e1.record(stream)
        kernel_events.append((k, e0, e1))
        if rep is not None:
            rep.sync()
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
    form_id = int(lib.qe_fused_form(algo.handle))
    fused_form = FORM_NAMES.get(form_id, str(form_id))
    buf = (C.c_uint64 * 48)()
    m = lib.qe_fused_phase_ns(algo.handle, buf, 48)
    phases = None
    if m >= 4:
        ks = (m - 1) // 3
        names = {3: ("select_env_step_us", "target_pipeline_us", "commit_and_sort_us"), 5: ("select_step_targets_us", "commit_us", "sort_us")}.get(
            form_id, ("select_step_register_us", "td_first_pass_us", "td_deferred_us"))
        phases = {name: sum(buf[1 + ph + 3 * j] - buf[ph + 3 * j] for j in range(ks)) / ks / 1e3 for ph, name in enumerate(names)}
        if form_id == 5:
            phases["sort_scatter_us"] = sum(buf[32 + j] - buf[2 + 3 * j] for j in range(ks)) / ks / 1e3
            phases["sort_buckets_us"] = sum(buf[3 + 3 * j] - buf[32 + j] for j in range(ks)) / ks / 1e3
        if form_id == 3:
            phases["commit_us"] = sum(buf[32 + j] - buf[2 + 3 * j] for j in range(ks)) / ks / 1e3
            phases["sort_us"] = sum(buf[3 + 3 * j] - buf[32 + j] for j in range(ks)) / ks / 1e3
    gather_peak = float(lib.qe_debug_gather_gbs(algo.handle)) if rank == 0 else None
    # The workload drifts: a greedy policy on a deterministic MDP herds agents onto the same rows, and the rows get more
    # crowded as the table is learned.  value_long: a fresh run, vector steps 0..512, one number; late_training: its window
    # of steps 256..288.
    late = None
    value_long = None
    if world == 1 and workload not in SMALL and not args.no_late:
        del algo, env
        algo, env = make()
        ep_ret = torch.zeros(n, dtype=torch.float32, device=dev)
        ag = env.agents_struct(ep_ret)
        t_next[0] = 0
        launch(SYNC_EVERY)  # (module / scratch already warm; these steps are part of the horizon but not of the clock)
        horizon = 512
        marks = [torch.cuda.Event(enable_timing=True)]
        marks[0].record(stream)
        while t_next[0] < horizon:
            for _ in range(4):
                launch(SYNC_EVERY)
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            marks.append(ev)
        torch.cuda.synchronize()
        capi.check(lib.qe_sync(algo.handle, C.c_void_p(stream.cuda_stream)))
        seg_ms = [marks[j].elapsed_time(marks[j + 1]) / 32 for j in range(len(marks) - 1)]  # ms per vector step, 32-step windows from step 8
        long_ms = marks[0].elapsed_time(marks[-1])
        long_steps = 32 * (len(marks) - 1)
        value_long = {"value": n * long_steps / (long_ms * 1e-3), "unit": "agent-steps/s", "vector_steps": f"8..{8 + long_steps} of a fresh run",
                      "ms_per_step": long_ms / long_steps, "roofline_frac": n * alg_bytes(a) * long_steps / (long_ms * 1e-3) / 1e9 / measured_peak()[0],
                      "ms_per_step_by_32_step_window": [round(x, 4) for x in seg_ms]}
        jl = min(len(seg_ms) - 1, (256 - 8) // 32 + 1)
        late = {"td_update_form": FORM_NAMES.get(int(lib.qe_fused_form(algo.handle)), "?"), "after_vector_steps": 8 + 32 * jl, "steps": 32, "ms_per_step": seg_ms[jl],
                "value": n / (seg_ms[jl] * 1e-3), "unit": "agent-steps/s", "roofline_frac": n * alg_bytes(a) / (seg_ms[jl] * 1e-3) / 1e9 / measured_peak()[0]}
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
                   "roofline_frac": n * alg_bytes(a) / (acc_ms * 1e-3) / 1e9 / peak_gbs, "kernel": "fused_kernel<MDP,2,ACC>" if workload not in SMALL else "fused_kernel<TTT,2,ACC>"}
        del algo, env

    # ---------------- e2e through the public API: per step H2D of that step's uniforms (pinned) + D2H of the results
    algo, env = make()
    rt = SingleThreadQLearning(algo, ConstantSchedule(LR), ConstantSchedule(EPS))
    rt.history_mode = "summary"
    runner = D.ReplicatedQLearning(rt, tp, sync_every=SYNC_EVERY, carry_over=True) if tp is not None else rt
    # config 2's 128 agents take ~10 us per vector step: one host round trip per step would measure the host.  The caller
    # asks for 64 steps per call there (their uniforms go down in one copy, their results come back once), one elsewhere.
    call_steps = 64 if (workload in SMALL and world == 1) else 1
    Ke = (min(K, 2048) // call_steps) * call_steps if call_steps > 1 else min(K, 20)
    Ke = max(Ke, call_steps)
    We = max(call_steps, min(W, 4 * call_steps)) if call_steps > 1 else W
    slots = env.slots
    u_host = torch.empty((We + Ke, n, slots), dtype=torch.int32).pin_memory()
    u_host.numpy().view(np.uint32)[:] = draw_uniforms(STREAM_SEED, 0, We + Ke, n, slots, agent0=env.agent0)
    pre = PredrawnUniforms(u_host.numpy().view(np.uint32))  # no copy: already contiguous uint32 (pinned)
    algo._rng = env._rng = pre
    sd = {"states": None, "infos": {}, "rewards": np.zeros(n, dtype=np.float32)}
    for _ in range(We // call_steps):
        _, _, _, sd = runner.run_steps(call_steps, env, sd)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(Ke // call_steps):
        _, _, _, sd = runner.run_steps(call_steps, env, sd)  # returns host copies of the agents' running returns
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = n * slots * 4 + n * 4 + 12  # the step's uniforms + the agents' running returns (state dict) + eps/lr of the step
    d2h = n * 4 + 16                  # the running returns + {sum, count} of the episodes that finished in the step
    e2e = {"value": world * n * Ke / e2e_s, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
           "steps": Ke, "api": (f"ReplicatedQLearning(sync_every={SYNC_EVERY}, carry_over=True)." if tp is not None else "SingleThreadQLearning.") +
           f"run_steps({call_steps}, env, state_dict): the step's pre-drawn uniforms (PredrawnUniforms, pinned host memory) and the state dict's running returns go host->device, the returns and the episode statistics come back, every step"}
    del algo, env, runner, rt

    # The same call with the engine's own counter stream instead of pre-drawn uniforms: the only per-step host traffic left is
    # the state dictionary's running returns (4 bytes per agent each way) -- what a caller that does not inject a random
    # stream gets.  Reported beside `e2e`, never instead of it (N = 1 only; a failure here must not cost the line).
    e2e_engine_rng = None
    if world == 1:
        try:
            algo, env = make()
            rt = SingleThreadQLearning(algo, ConstantSchedule(LR), ConstantSchedule(EPS))
            rt.history_mode = "summary"
            sd = {"states": None, "infos": {}, "rewards": np.zeros(n, dtype=np.float32)}
            for _ in range(max(1, We // call_steps)):
                _, _, _, sd = rt.run_steps(call_steps, env, sd)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(Ke // call_steps):
                _, _, _, sd = rt.run_steps(call_steps, env, sd)
            torch.cuda.synchronize()
            es = time.perf_counter() - t0
            e2e_engine_rng = {"value": n * Ke / es, "unit": "agent-steps/s", "h2d_bytes_per_step": n * 4 + 12, "d2h_bytes_per_step": n * 4 + 16, "steps": Ke,
                              "api": f"SingleThreadQLearning.run_steps({call_steps}, env, state_dict) with the engine's counter stream: only the state dict's running returns travel, both ways, every step"}
            del algo, env, rt
        except Exception as exc:  # noqa: BLE001
            e2e_engine_rng = {"value": None, "error": repr(exc)[:200]}

    # ---------------- sharded 100M-state table (config 4): peer memory over NVLink, one persistent kernel per GPU
    sharded = None
    parity = None
    if tp is not None and not args.no_sharded:
        s4, a4, n4, desc4 = WORKLOADS["c4"]
        sh = D.ShardedQLearning(s4, a4, GAMMA, n4, tp, env_seed=ENV_SEED, p_term=P_TERM, seed=STREAM_SEED, device=local)
        sh.fill_random(TABLE_SEED)
        sh.reset()
        eps_s, lr_s = ConstantSchedule(EPS), ConstantSchedule(LR)
        sh.run_steps(8, eps_s, lr_s)
        sh.sync()
        sync_all()
        ks = 16
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        sh.run_steps(ks, eps_s, lr_s)
        b1.record(stream)
        sh.sync()
        sync_all()
        ms = max_over_ranks(b0.elapsed_time(b1))
        ph = sh.phase_us()
        cs = sh.table_checksum()
        steps_done = 8 + ks
        sh.close()
        del sh
        # the same number of steps of the same job on ONE GPU (rank 0 alone, same engine with a world of one): equal tables?
        cs1 = None
        ms1 = None
        if rank == 0:
            class _Solo(D.Transport):
                rank, world_size = 0, 1

                def all_reduce_sum_(self, t):
                    return t

                def all_gather_rows(self, t):
                    return t

                def barrier(self):
                    pass

            solo = D.ShardedQLearning(s4, a4, GAMMA, n4, _Solo(), env_seed=ENV_SEED, p_term=P_TERM, seed=STREAM_SEED, device=local)
            solo.fill_random(TABLE_SEED)
            solo.reset()
            solo.run_steps(8, eps_s, lr_s)
            solo.sync()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            solo.run_steps(ks, eps_s, lr_s)
            c1.record(stream)
            solo.sync()
            ms1 = c0.elapsed_time(c1)
            cs1 = solo.table_checksum()
            solo.close()
            del solo
        sync_all()
        sharded = {"workload": f"c4: {desc4}", "value": n4 * ks / (ms * 1e-3), "unit": "agent-steps/s", "scaling": "strong", "steps": ks,
                   "ms_per_step": ms / ks, "states": s4, "actions": a4, "agents": n4, "phase_us_per_step_rank0": ph,
                   "table_checksum": f"{cs:016x}", "table_checksum_single_gpu": None if cs1 is None else f"{cs1:016x}",
                   "equals_single_gpu_table": None if cs1 is None else bool(cs == cs1), "vector_steps_compared": steps_done,
                   "single_gpu_same_engine": None if ms1 is None else {"value": n4 * ks / (ms1 * 1e-3), "ms_per_step": ms1 / ks},
                   "exchange": "no collective on the data path: rows, writer records and targets are peer loads / stores over NVLink (slabs in "
                               "torch symmetric memory), the per-step order is a distributed stable sort, 3 flag barriers in peer memory per step"}
        # ---------------- does the multi-GPU path compute the right thing HERE?  Small cases against the C oracle on rank 0.
        parity = multi_gpu_parity(tp, local, torch)

    if rank != 0:
        return None
    # ---------------- roofline + CPU baseline (rank 0)
    peak, peak_src = measured_peak()
    balg = alg_bytes(a)
    per_launch_ms = kernel_ms / len(kernel_events)
    steps_per_launch = K / len(kernel_events)
    achieved = n * steps_per_launch * balg / (per_launch_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel: from the committed ncu capture of this command (never measured under the bench)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            ent = tj.get(f"{workload}_form{form_id}")
            if isinstance(ent, dict):
                traffic, traffic_src = ent.get("dram_bytes_per_launch"), ent.get("source")
        except Exception:  # noqa: BLE001
            traffic = None
    kname = FORM_KERNELS.get(form_id, "fused_kernel") + (("<TTT>" if form_id == 4 else "<TTT,2>") if workload in SMALL else ("<MDP,1>" if a <= 8 else "<MDP,2>"))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src, "kernel": kname, "algorithmic_bytes_per_agent_step": balg,
                "algorithmic_bytes_per_launch": n * steps_per_launch * balg, "avg_launch_ms": per_launch_ms,
                "steps_per_launch": steps_per_launch, "phase_us_per_step": phases,
                "gather_peak": {"value": gather_peak, "unit": "GB/s", "what": "dependency-free random whole-row gathers over this table (qe_debug_gather_gbs): "
                                "the ceiling of the engine's dominant access pattern, to be read beside the copy peak"}}
    cpu_baseline = None
    if world == 1:
        base = cpu_baselines_subprocess(workload)
        if base is not None and "single_thread" in base:
            cpu_baseline = {"value": base["single_thread"]["value"], "unit": "agent-steps/s", "cores": 1, "kind": "reference",
                            "sample": base["single_thread"]["sample"], **{k: v for k, v in base.items() if k != "single_thread"}}
        elif base is not None:
            cpu_baseline = {"value": base["c_port"]["value"], "unit": "agent-steps/s", "cores": base["c_port"]["cores"], "kind": "port",
                            "sample": base["c_port"]["sample"], **{k: v for k, v in base.items() if k != "c_port"}}
    cfg = {"workload": f"{workload}: {desc}" + (f", one replica per GPU, Q-delta all-reduce every {SYNC_EVERY} steps (BASELINE config 5)" if world > 1 else ""),
           "states": s, "actions": a, "agents_per_gpu": n, "agents": n * world, "eps": EPS, "lr": LR, "gamma": GAMMA, "p_term": P_TERM,
           "table_init": "uniform[0,1)" if workload not in SMALL else "zeros", "rng": "on-device counter stream",
           "steps_per_launch": per_launch,
           "timing": "CUDA events around the K timed steps (max over ranks); no L2 flush: every vector step touches the dense table (64 MB at "
                     "config 3) plus ~90 MB of per-agent / per-position arrays, more than the 126 MB L2 (ncu: L2 hit rate 65 %)",
           "grid_blocks": grid_blocks, "td_update_form": fused_form, "timed_window": f"vector steps {W}..{W + K} of the run",
           "warmup_requested": args.warmup, "warmup_used": W, "late_training": late}
    out = {
        "metric": "agent-steps/s", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg, "roofline": roofline, "e2e": e2e, "e2e_engine_rng": e2e_engine_rng, "gpu_launches": gpu_launches, "clocks": clocks,
        "episodes": episodes,
    }
    if value_long is not None:
        out["value_long"] = value_long
    if parity is not None:
        out["multi_gpu_parity"] = parity.get("result", "FAIL")
        out["multi_gpu_parity_detail"] = parity
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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "_cpu"])
    ap.add_argument("--workload", default=None, choices=[None, *WORKLOADS])
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the bounded run of the sharded 100M-state table")
    ap.add_argument("--no-late", action="store_true", help="N = 1: skip the extra measurement after 256 vector steps")
    args = ap.parse_args()
    if args.impl == "_cpu":
        out = cpu_baselines(args.workload or "c3", quick=os.environ.get("BENCH_QUICK_CPU", "0") == "1")
    elif args.impl == "reference":
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
