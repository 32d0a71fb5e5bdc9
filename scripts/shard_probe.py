"""Random-access throughput over the local and a peer's table shard as the sharded engine maps them (CUDA IPC), under torchrun."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dist_classicrl_b200 import capi, distributed as D
local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tp = D.TorchDistTransport()
sh = D.ShardedQLearning(100_000_000, 8, 0.99, 1 << 22, tp, device=local)
sh.fill_random(1)
torch.cuda.synchronize(); dist.barrier()
lib = capi.lib()
for r in range(world):
    if tp.rank == r:   # one rank at a time: the other GPU is idle
        for peer in range(world):
            print(f"rank {r} -> shard of rank {peer}: loads {lib.qe_shard_probe(sh._h, peer, 0):.2f} G/s, stores {lib.qe_shard_probe(sh._h, peer, 1):.2f} G/s", flush=True)
    dist.barrier()
# both at once (bidirectional)
peer = (tp.rank + 1) % world
print(f"simultaneous: rank {tp.rank} -> {peer}: loads {lib.qe_shard_probe(sh._h, peer, 0):.2f} G/s", flush=True)
dist.barrier()
sh.close(); dist.destroy_process_group()
