"""Timing of the peer-memory sharded table (config 4 by default), one rank per GPU under torchrun, or a single process.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/shard_bench.py [S A N steps]
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dist_classicrl_b200 import distributed as D  # noqa: E402
from dist_classicrl_b200.schedules import ConstantSchedule  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tp = D.TorchDistTransport()
else:
    class _Solo(D.Transport):
        rank, world_size = 0, 1
        def all_reduce_sum_(self, t): return t
        def all_gather_rows(self, t): return t
        def barrier(self): pass
    tp = _Solo()
S, A, N, steps = 100_000_000, 8, 1 << 22, 16
if len(sys.argv) > 4:
    S, A, N, steps = int(float(sys.argv[1])), int(sys.argv[2]), int(float(sys.argv[3])), int(sys.argv[4])
sh = D.ShardedQLearning(S, A, 0.99, N, tp, env_seed=0, p_term=0.05, seed=0, device=local)
sh.fill_random(1)
sh.reset()
e, l = ConstantSchedule(0.1), ConstantSchedule(0.1)
sh.run_steps(8, e, l)
sh.sync()
tp.barrier()
best = None
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    sh.run_steps(steps, e, l)
    b.record()
    sh.sync()
    ms = torch.tensor([a.elapsed_time(b)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    best = ms if best is None else min(best, ms)
    if tp.rank == 0:
        print(f"rep {rep}: {ms / steps * 1e3:.1f} us/step, {N * steps / ms / 1e6:.3f} G agent-steps/s", flush=True)
ph = sh.phase_us()
if world > 1:
    allph = [None] * world
    dist.all_gather_object(allph, ph)
else:
    allph = [ph]
if tp.rank == 0:
    for g, x in enumerate(allph):
        print(f"rank {g} phases us:", {k: round(v, 1) for k, v in x.items()}, flush=True)
cs = sh.table_checksum()
if tp.rank == 0:
    print(json.dumps({"sharded": True, "world": world, "states": S, "actions": A, "agents": N, "steps": steps, "ms_per_step": best / steps,
                      "value": N * steps / (best * 1e-3), "unit": "agent-steps/s", "table_checksum": f"{cs:016x}", "vector_steps_done": 8 + 3 * steps}), flush=True)
sh.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
