"""Development aid (torchrun, one rank per GPU): where the time of one replicated-table sync goes."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dist_classicrl_b200 import capi  # noqa: E402
from dist_classicrl_b200 import distributed as D  # noqa: E402
from dist_classicrl_b200.algorithms.base_algorithms.q_learning_optimal import OptimalQLearningBase  # noqa: E402
from dist_classicrl_b200.algorithms.runtime import SingleThreadQLearning  # noqa: E402
from dist_classicrl_b200.schedules import ConstantSchedule  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tp = D.TorchDistTransport()
algo = OptimalQLearningBase(1_000_000, 16, 0.99, seed=0, device=local)
algo.fill_random(1)
rep = D.ReplicatedQLearning(SingleThreadQLearning(algo, ConstantSchedule(0.1), ConstantSchedule(0.1)), tp, sync_every=8)
lib, h = capi.lib(), algo.handle
st = torch.cuda.current_stream()
sp = lambda: __import__("ctypes").c_void_p(st.cuda_stream)  # noqa: E731
for name, fn in (("delta", lambda: capi.check(lib.qe_table_delta_dense(h, rep.base.data_ptr(), rep.delta.data_ptr(), sp()))),
                 ("all_reduce 64 MB", lambda: tp.all_reduce_sum_(rep.delta)),
                 ("merge", lambda: capi.check(lib.qe_table_merge_dense(h, rep.base.data_ptr(), rep.delta.data_ptr(), sp()))),
                 ("sync()", rep.sync)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record()
    torch.cuda.synchronize()
    if dist.get_rank() == 0:
        print(f"{name:18s} {a.elapsed_time(b) / 10:.3f} ms")
dist.barrier()
dist.destroy_process_group()
