"""World-size-2 (and 3) tests of the multi-GPU host logic on CPU: gloo process group + the in-process loopback
transport.  Covers the state-range partition, stable routing to owners and back (the order-preservation the exact
sharded TD update relies on), migration merge by global id and the replicated-table delta rule."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from dist_classicrl_b200 import distributed as D  # noqa: E402


def test_partition_covers_every_state_once():
    for s, g in ((10, 3), (100_000_000, 8), (7, 8), (19683, 2)):
        ranges = [D.shard_range(s, g, r) for r in range(g)]
        assert ranges[0][0] == 0 and ranges[-1][1] == s
        for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
            assert a1 == b0 and a0 <= a1
        states = np.unique(np.concatenate([np.arange(0, min(s, 1000)), np.arange(max(s - 1000, 0), s)]))
        own = D.owner_of(states, s, g)
        for st, o in zip(states[::97], own[::97]):
            lo, hi = ranges[int(o)]
            assert lo <= st < hi


def _routing_body(tp):
    g, r = tp.world_size, tp.rank
    rng = np.random.default_rng(100 + r)
    n = 50 + 7 * r
    gid = torch.from_numpy(np.sort(rng.choice(10_000, n, replace=False)).astype(np.int32)) * g + r  # distinct across ranks
    state = torch.from_numpy(rng.integers(0, 1000, n).astype(np.int32))
    owner = D.owner_of(state, 1000, g).to(torch.int64)
    rows = torch.stack([gid, state], dim=1)
    got, order, in_counts = D.route(tp, rows, owner)
    lo, hi = D.shard_range(1000, g, r)
    assert bool(((got[:, 1] >= lo) & (got[:, 1] < hi)).all())  # everything that arrives is ours
    # inside each source block the global ids are still ascending (stable bucketing)
    off = 0
    for c in in_counts:
        blk = got[off:off + c, 0]
        assert bool((blk[1:] > blk[:-1]).all())
        off += c
    # answers find their way back to the asking row
    answers = (got[:, 0:1] * 3 + got[:, 1:2]).contiguous()
    back = D.route_back(tp, answers, in_counts, order)
    assert torch.equal(back.reshape(-1), gid * 3 + state)
    # replicated-table rule
    base = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    local = base + float(r + 1)
    merged = D.merge_deltas(tp, local, base)
    assert torch.equal(merged, base + float(sum(range(1, g + 1))))
    assert tp.all_reduce_max_int(r) == g - 1
    allrows = tp.all_gather_rows(rows)
    assert allrows.shape[0] == sum(50 + 7 * k for k in range(g))
    return int(got.shape[0])


@pytest.mark.parametrize("world", [2, 3])
def test_loopback_transport_routing(world):
    counts = D.run_loopback(world, _routing_body)
    assert sum(counts) == sum(50 + 7 * k for k in range(world))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, ret):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = _routing_body(D.TorchDistTransport())
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2_routing():
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gloo_worker, args=(world, port, ret), nprocs=world, join=True)
        assert sum(ret.values()) == sum(50 + 7 * k for k in range(world))
