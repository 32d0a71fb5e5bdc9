"""BASELINE config 5 sweep (torchrun, one rank per GPU): replicated table, hash MDP 1M x 16, agents per GPU 128 .. 2^20, merge
period K in {1, 8, 64}.  Device-timed (max over ranks), `steps` vector steps per point after a warm-up of the same length;
prints one JSON line per point on rank 0 and writes gpurun_out/c5_sweep_n<world>.json.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/c5_sweep.py [steps]
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dist_classicrl_b200 import capi  # noqa: E402
from dist_classicrl_b200 import distributed as D  # noqa: E402
from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase  # noqa: E402
from dist_classicrl_b200.algorithms.runtime import SingleThreadQLearning  # noqa: E402
from dist_classicrl_b200.environments import HashMDPVecEnv  # noqa: E402
from dist_classicrl_b200.rng import explore_threshold  # noqa: E402
from dist_classicrl_b200.schedules import ConstantSchedule  # noqa: E402

S, A, EPS, LR = 1_000_000, 16, 0.1, 0.1
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 64
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
tp = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    tp = D.TorchDistTransport()
lib = capi.lib()
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
out = []
for n in (128, 1024, 8192, 65536, 1 << 20):
    for K in (1, 8, 64):
        algo = OptimalQLearningBase(S, A, 0.99, seed=0, device=local)
        algo.fill_random(1)
        env = HashMDPVecEnv(n, S, A, env_seed=0, p_term=0.05, seed=0, device=local, output="torch")
        env.agent0 = rank * n
        env.attach(algo)
        env.reset()
        ep_ret = torch.zeros(n, dtype=torch.float32, device=dev)
        ag = env.agents_struct(ep_ret)
        rep = D.ReplicatedQLearning(SingleThreadQLearning(algo, ConstantSchedule(LR), ConstantSchedule(EPS)), tp, sync_every=K) if tp is not None else None
        t = [0]

        def launch(k):
            th = np.full(k, explore_threshold(EPS), dtype=np.uint64)
            lrs = np.full(k, LR, dtype=np.float32)
            run = capi.QeRun()
            run.steps = k
            run.explore_thresholds_host = th.ctypes.data_as(C.c_void_p)
            run.learning_rates_host = lrs.ctypes.data_as(C.c_void_p)
            run.slots = env.slots
            run.stream_seed = run.env_stream_seed = 0
            run.t0 = run.env_t0 = t[0]
            run.agent0 = env.agent0
            run.use_masks = 1
            run.empty_all = 1
            capi.check(lib.qe_fused_steps(algo.handle, C.byref(ag), C.byref(run), sp))
            t[0] += k

        def window(total):
            done = 0
            while done < total:
                k = min(K, total - done)
                launch(k)
                if rep is not None:
                    rep.sync()
                done += k

        window(steps)
        capi.check(lib.qe_sync(algo.handle, sp))
        if tp is not None:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        window(steps)
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if tp is not None:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        rec = {"gpus": world, "agents_per_gpu": n, "merge_every": K, "steps": steps, "us_per_step": ms / steps * 1e3,
               "value": world * n * steps / (ms * 1e-3), "unit": "agent-steps/s", "form": int(lib.qe_fused_form(algo.handle))}
        out.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
        del rep, env, algo, ag, ep_ret
        torch.cuda.empty_cache()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/c5_sweep_n{world}.json", "w"), indent=1)
if tp is not None:
    dist.barrier()
    dist.destroy_process_group()
