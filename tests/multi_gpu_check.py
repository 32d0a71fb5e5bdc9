"""Sharded / replicated modes over NCCL, one rank per GPU, checked against the C oracle (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dist_classicrl_b200 import distributed as D  # noqa: E402
from dist_classicrl_b200.schedules import ConstantSchedule  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import rng as orng  # noqa: E402
from oracle.envs import T_INIT  # noqa: E402

EPS, LR, GAMMA, P_TERM = 0.1, 0.1, 0.99, 0.05


def random_table(S, A, seed):
    x = (np.arange(S * A, dtype=np.uint64) ^ np.uint64((seed * 0x9E3779B9) & 0xFFFFFFFF)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return ((x >> np.uint64(8)).astype(np.float32) * np.float32(2.0**-24)).reshape(S, A)


def main() -> int:
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tp = D.TorchDistTransport()
    ok = True
    for S, A, N, steps in ((5000, 8, 6000, 8), (300_000, 16, 100_000, 6)):
        seed, env_seed, table_seed = 7, 3, 1
        sh = D.ShardedQLearning(S, A, GAMMA, N, tp, env_seed=env_seed, p_term=P_TERM, seed=seed, device=local)
        sh.fill_random(table_seed)
        sh.reset()
        sh.run_steps(steps, ConstantSchedule(EPS), ConstantSchedule(LR))
        table = sh.gather_table()
        states, rets = sh.gather_agents()
        if tp.rank == 0:
            tt = int(math.ceil(P_TERM * 2.0**32))
            st_o, mk_o = co.mdp_reset(orng.draw_uniforms(seed, T_INIT, 1, N, 4)[0], S, A, env_seed)
            q_o = random_table(S, A, table_seed)
            rew = np.zeros(N, dtype=np.float32)
            th = np.full(steps, orng.explore_threshold(EPS), dtype=np.uint64)
            lr = np.full(steps, LR, dtype=np.float32)
            res = co.run(co.ENV_MDP, q_o, None, st_o, mk_o, num_states=S, env_seed=env_seed, term_thresh=tt, uniforms=None, slots=4,
                         stream_seed=seed, steps=steps, eps_thresh=th, lr=lr, gamma=GAMMA, empty_all=A > 10, agent_rewards=rew)
            good = res["rc"] == 0 and np.array_equal(states, st_o) and np.array_equal(table, q_o) and np.array_equal(rets, rew)
            print(f"sharded G={tp.world_size} S={S} A={A} N={N} steps={steps}: "
                  f"states_equal={np.array_equal(states, st_o)} table_equal={np.array_equal(table, q_o)} returns_equal={np.array_equal(rets, rew)}",
                  flush=True)
            ok = ok and good
        sh.close()
        del sh
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if tp.rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
