"""Per-section wall times of the sharded step (diagnostics; run under torchrun)."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dist_classicrl_b200 import distributed as D
from dist_classicrl_b200.schedules import ConstantSchedule
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tp = D.TorchDistTransport()
S, A, N = 100_000_000, 8, 1 << 22
sh = D.ShardedQLearning(S, A, 0.99, N, tp, env_seed=0, p_term=0.05, seed=0, device=local)
sh.fill_random(1); sh.reset()
e, l = ConstantSchedule(0.1), ConstantSchedule(0.1)
sh.run_steps(3, e, l)
sh.profile = {}
r0 = sh.rounds_total
sh.run_steps(5, e, l)
if tp.rank == 0:
    tot = sum(sh.profile.values())
    print(f"rounds/step {(sh.rounds_total - r0) / 5:.1f}; total {tot / 5 * 1e3:.2f} ms/step")
    for k, v in sorted(sh.profile.items(), key=lambda x: -x[1]):
        print(f"  {k:18s} {v / 5 * 1e3:8.3f} ms/step  {100 * v / tot:5.1f}%")
dist.barrier(); dist.destroy_process_group()
