"""Multi-GPU modes of the engine: one process per GPU, ``torch.distributed`` for the plumbing.

The reference's distributed trainer (``algorithms/runtime/q_learning_async_dist.py``, MPI:59-447) keeps ONE table on
rank 0 and ships transitions to it; both modes below are new relative to it and keep its entry-point shape
(``train`` / ``run_steps`` on a runtime object that owns an algorithm and two schedules).

* :class:`ShardedQLearning` (BASELINE config 4) -- the table is split by state range, ``owner(s) = s // ceil(S/G)``;
  an agent lives on the rank that owns its current state, so select and the environment step are local.  The TD
  update of agent ``i`` writes ``Q[s_i, a_i]`` (local) and bootstraps from row ``s'_i`` (anywhere).  Exactness
  (results identical to the single-GPU engine, i.e. to the reference's sequential loop in global agent order) is
  reached by a fixed-point iteration over the *remote* bootstrap values only:

      1. requests ``(global id, s')`` travel to ``owner(s')`` once (all-to-all);
      2. every rank runs the exact sequential update on its own agents with the current remote values held fixed
         (``qe_learn`` in hold mode: values are published, the table stays untouched), then answers the requests it
         received from the published values of the earlier local writers of each row (``qe_serve_bootstrap``);
      3. the answers travel back (all-to-all); when no answer changed anywhere the values are exact -- the
         equations are triangular in the global agent index, so the fixed point is unique -- and every rank commits.

  The number of rounds is the number of cross-shard hops on the longest dependency chain of the step (2-4 at the
  densities of config 4).  Then every agent migrates to ``owner(s')`` (all-to-all) and the arrivals are merged in
  global-id order, which makes the local agent order the global one.
* :class:`ReplicatedQLearning` (BASELINE config 5) -- every rank holds the whole table and ``N/G`` agents and runs
  the fused loop locally; every ``sync_every`` vector steps ``Q <- Q_base + sum_g (Q_g - Q_base)`` with one
  all-reduce of the deltas (``qe_table_delta_dense`` -> all-reduce -> ``qe_table_merge_dense``).

Transports: :class:`TorchDistTransport` (NCCL on GPUs, gloo in the CPU tests) and :class:`LoopbackTransport`
(``G`` virtual ranks as threads of one process sharing one GPU -- parity tests of the multi-rank logic on one device).
"""

from __future__ import annotations

import os

import ctypes as C
import math
import threading
from typing import Callable, Sequence

import numpy as np

from dist_classicrl_b200 import capi
from dist_classicrl_b200.rng import T_INIT, explore_threshold


def _torch():
    import torch

    return torch


# ----------------------------------------------------------------------------------------------------- partition
def shard_rows(num_states: int, world_size: int) -> int:
    """Rows per shard: ``ceil(S / G)``."""
    return -(-int(num_states) // int(world_size))


def shard_range(num_states: int, world_size: int, rank: int) -> tuple[int, int]:
    b = shard_rows(num_states, world_size)
    lo = min(rank * b, num_states)
    return lo, min(lo + b, num_states)


def owner_of(states, num_states: int, world_size: int):
    """Owning rank of every state id (tensor or array in, same kind out)."""
    return states // shard_rows(num_states, world_size)


# ----------------------------------------------------------------------------------------------------- transports
class Transport:
    """What the multi-GPU modes need from the fabric.  Tensors are 2-D ``[n, cols]`` int32 (floats travel as bits)."""

    rank: int
    world_size: int

    def all_to_all_v(self, send: Sequence):  # send[d] -> rank d; returns recv[s] from rank s
        raise NotImplementedError

    def all_to_all_known(self, send: Sequence, in_counts: Sequence[int]):
        """Like :meth:`all_to_all_v` when the receive counts are already known (no count exchange, no host sync)."""
        return self.all_to_all_v(send)

    def all_reduce_sum_(self, tensor):
        raise NotImplementedError

    def all_reduce_max_int(self, value: int) -> int:
        raise NotImplementedError

    def all_reduce_max_flag(self, flag) -> int:
        """Max over ranks of a one-element integer tensor that already lives on the transport's device."""
        return self.all_reduce_max_int(int(flag.item()))

    def all_gather_rows(self, tensor):  # concatenation over ranks along dim 0 (row counts may differ)
        raise NotImplementedError

    def barrier(self) -> None:
        raise NotImplementedError


class TorchDistTransport(Transport):
    """``torch.distributed`` process group (backend nccl: NVLink/NVSwitch; gloo: CPU tests)."""

    def __init__(self, group=None) -> None:
        import torch.distributed as dist

        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)

    def all_to_all_v(self, send):
        torch, dist = _torch(), self._dist
        g = self.world_size
        dev = send[0].device
        cols = send[0].shape[1]
        counts = torch.tensor([int(x.shape[0]) for x in send], dtype=torch.int64, device=dev)
        incoming = torch.empty_like(counts)
        dist.all_to_all_single(incoming, counts, group=self.group)
        in_splits = incoming.tolist()
        out_splits = counts.tolist()
        flat = torch.cat(list(send), dim=0).contiguous()
        recv = torch.empty((sum(in_splits), cols), dtype=flat.dtype, device=dev)
        if dist.get_backend(self.group) == "gloo":  # gloo has no all_to_all_single with splits on every build
            outs = [torch.empty((in_splits[s], cols), dtype=flat.dtype, device=dev) for s in range(g)]
            reqs = []
            for peer in range(g):
                if peer == self.rank:
                    outs[peer].copy_(send[peer])
                    continue
                reqs.append(dist.isend(send[peer].contiguous(), dist.get_global_rank(self.group, peer) if self.group else peer, group=self.group))
                reqs.append(dist.irecv(outs[peer], dist.get_global_rank(self.group, peer) if self.group else peer, group=self.group))
            for r in reqs:
                r.wait()
            return outs
        dist.all_to_all_single(recv, flat, [n for n in in_splits], [n for n in out_splits], group=self.group)
        return list(torch.split(recv, in_splits, dim=0))

    def all_to_all_known(self, send, in_counts):
        torch, dist = _torch(), self._dist
        if dist.get_backend(self.group) == "gloo":
            return self.all_to_all_v(send)
        flat = torch.cat(list(send), dim=0).contiguous()
        recv = torch.empty((sum(in_counts), flat.shape[1]), dtype=flat.dtype, device=flat.device)
        dist.all_to_all_single(recv, flat, list(in_counts), [int(x.shape[0]) for x in send], group=self.group)
        return list(torch.split(recv, list(in_counts), dim=0))

    def all_reduce_max_flag(self, flag) -> int:
        if self._dist.get_backend(self.group) == "gloo":
            flag = flag.cpu()
        flag = flag.clone()
        self._dist.all_reduce(flag, op=self._dist.ReduceOp.MAX, group=self.group)
        return int(flag.item())

    def all_reduce_sum_(self, tensor):
        self._dist.all_reduce(tensor, op=self._dist.ReduceOp.SUM, group=self.group)
        return tensor

    def all_reduce_max_int(self, value: int) -> int:
        torch = _torch()
        dev = "cuda" if self._dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def all_gather_rows(self, tensor):
        torch = _torch()
        n = torch.tensor([tensor.shape[0]], dtype=torch.int64, device=tensor.device)
        sizes = [torch.empty_like(n) for _ in range(self.world_size)]
        self._dist.all_gather(sizes, n, group=self.group)
        sizes = [int(x.item()) for x in sizes]
        cap = max(sizes) if sizes else 0
        pad = torch.zeros((cap,) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
        pad[: tensor.shape[0]] = tensor
        outs = [torch.empty_like(pad) for _ in range(self.world_size)]
        self._dist.all_gather(outs, pad, group=self.group)
        return torch.cat([o[:k] for o, k in zip(outs, sizes)], dim=0)

    def barrier(self) -> None:
        self._dist.barrier(group=self.group)


class _LoopbackHub:
    def __init__(self, world_size: int) -> None:
        self.world_size = world_size
        self.bar = threading.Barrier(world_size)
        self.mail: list = [None] * world_size


class LoopbackTransport(Transport):
    """``G`` virtual ranks = ``G`` threads of this process (see :func:`run_loopback`)."""

    def __init__(self, hub: _LoopbackHub, rank: int) -> None:
        self.hub, self.rank, self.world_size = hub, rank, hub.world_size

    def _exchange(self, item):
        hub = self.hub
        hub.mail[self.rank] = item
        hub.bar.wait()
        got = list(hub.mail)
        hub.bar.wait()
        return got

    def all_to_all_v(self, send):
        got = self._exchange(list(send))
        return [got[s][self.rank].clone() for s in range(self.world_size)]

    def all_reduce_sum_(self, tensor):
        got = self._exchange(tensor.clone())
        acc = got[0].clone()
        for s in range(1, self.world_size):  # fixed rank order: reproducible
            acc = acc + got[s]
        tensor.copy_(acc)
        return tensor

    def all_reduce_max_int(self, value: int) -> int:
        return max(self._exchange(int(value)))

    def all_gather_rows(self, tensor):
        return _torch().cat(self._exchange(tensor.clone()), dim=0)

    def barrier(self) -> None:
        self.hub.bar.wait()


def run_loopback(world_size: int, fn: Callable[[Transport], object]) -> list:
    """Run ``fn(transport)`` on ``world_size`` virtual ranks (threads); returns the per-rank results."""
    hub = _LoopbackHub(world_size)
    results: list = [None] * world_size
    errors: list = []

    def work(rank: int) -> None:
        try:
            results[rank] = fn(LoopbackTransport(hub, rank))
        except BaseException as exc:  # noqa: BLE001
            errors.append(exc)
            hub.bar.abort()

    threads = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world_size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


# ----------------------------------------------------------------------------------------------------- routing
def bucket_by_owner(owner, world_size: int):
    """Stable bucketing: ``(order, counts)`` with ``order`` = indices sorted by owner (input order kept inside a
    bucket, so sorted-by-global-id stays sorted) and ``counts[d]`` = items for rank ``d``."""
    torch = _torch()
    order = torch.sort(owner, stable=True).indices
    counts = torch.bincount(owner, minlength=world_size)[:world_size]
    return order, counts.tolist()


def route(transport: Transport, rows, owner):
    """Send the rows ``[n, cols]`` to their owners.  Returns ``(received rows in source-rank order, order, counts)``;
    ``order``/``counts`` let :func:`route_back` return per-row answers to where the rows came from."""
    torch = _torch()
    order, counts = bucket_by_owner(owner, transport.world_size)
    send = list(torch.split(rows[order], counts, dim=0))
    recv = transport.all_to_all_v(send)
    in_counts = [int(x.shape[0]) for x in recv]
    return (torch.cat(recv, dim=0) if recv else rows[:0]), order, in_counts


def route_back(transport: Transport, answers, in_counts, order, out_counts=None):
    """Inverse of :func:`route`: ``answers`` (one row per received row, same order) go back to the senders and are
    put into the senders' original row order.  ``out_counts`` (how many rows this rank sent to every owner) saves the
    count exchange."""
    torch = _torch()
    parts = list(torch.split(answers, in_counts, dim=0))
    back = transport.all_to_all_v(parts) if out_counts is None else transport.all_to_all_known(parts, out_counts)
    flat = torch.cat(back, dim=0)
    out = torch.empty_like(flat)
    out[order] = flat
    return out


def merge_deltas(transport: Transport, local, base):
    """Replicated-table rule: ``base + sum_g (local_g - base)`` (device-agnostic tensor form, used by the CPU tests
    and as the statement of what ``qe_table_delta_dense`` / ``qe_table_merge_dense`` compute)."""
    delta = local - base
    transport.all_reduce_sum_(delta)
    return base + delta


# ----------------------------------------------------------------------------------------------------- sharded table
class ShardedQLearning:
    """State-range-sharded Q-table over peer memory (BASELINE config 4; ``csrc/qe_shard.cuh``).

    Rank ``g`` owns the states ``[g * rows, (g + 1) * rows)``, ``rows = ceil(S / G)``, and for good the agents
    ``[g * ceil(N / G), ...)``.  One persistent kernel per GPU runs whole vector steps: rows, writer records and
    targets travel as peer loads / stores over NVLink, the per-step order is a distributed stable sort, phases are
    separated by flag barriers in peer memory -- no host synchronisation and no collective on the data path
    (``torch.distributed`` only carries the 64-byte CUDA IPC handles at construction and the gathers of the
    ``gather_*`` helpers).  With a :class:`LoopbackTransport` the ``G`` ranks share one GPU and run side by side in one
    cooperative launch (the one-GPU parity tests).

    Parameters mirror ``OptimalQLearningBase`` (QLO:84-90) plus the environment of ``HashMDPVecEnv``.  Results (table,
    agent states, episode returns) are identical to the single-GPU engine run on the same seeds.
    """

    launch_steps = 8  # vector steps per kernel launch

    def __init__(self, state_size: int, action_size: int, discount_factor: float, num_agents: int, transport: Transport,
                 env_seed: int = 0, p_term: float = 0.05, seed: int = 0, device: int | None = None) -> None:
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("ShardedQLearning needs a CUDA device: the engine has no CPU fallback")
        if state_size * action_size >= 2**32 or action_size > 32:
            raise ValueError("the hash MDP needs S*A < 2**32 and A <= 32")
        self.tp = transport
        self.rank, self.world = transport.rank, transport.world_size
        self.state_size, self.action_size, self.discount_factor = int(state_size), int(action_size), float(discount_factor)
        self.num_agents = int(num_agents)
        self.env_seed, self.term_threshold = int(env_seed), int(math.ceil(p_term * 2.0**32))
        self.seed = int(seed)
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device("cuda", self.device_index)
        self.lo, self.hi = shard_range(state_size, self.world, self.rank)
        self._lib = capi.lib()
        self._h = C.c_void_p()
        # The slab the peers read and write: torch symmetric memory (CUDA VMM, large pages, file descriptors exchanged by
        # torch) when the ranks are processes; CUDA IPC of a cudaMalloc block as the fallback (30x slower for random peer
        # accesses: QE_SHARD_IPC=1 forces it, for the record)
        self._symm = None
        slab = None
        if isinstance(transport, TorchDistTransport) and self.world > 1 and os.environ.get("QE_SHARD_IPC", "0") != "1":
            try:
                import torch.distributed._symmetric_memory as symm_mem

                nbytes = int(self._lib.qe_shard_slab_bytes(self.state_size, self.action_size, self.world, self.num_agents))
                buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.dev)
                import torch.distributed as dist

                hdl = symm_mem.rendezvous(buf, transport.group if transport.group is not None else dist.group.WORLD)
                self._symm = (buf, hdl)
                slab = C.c_void_p(int(buf.data_ptr()))
            except Exception as exc:  # noqa: BLE001
                import warnings

                warnings.warn(f"torch symmetric memory is unavailable ({exc!r}): the sharded table falls back to CUDA IPC mappings (slow peer access)")
                self._symm = None
        capi.check(self._lib.qe_shard_create(self.state_size, self.action_size, self.discount_factor, self.device_index, self.rank,
                                             self.world, self.num_agents, self.env_seed, slab, C.byref(self._h)))
        self.n_home = int(self._lib.qe_shard_info(self._h, 1))
        self.n_here = int(self._lib.qe_shard_info(self._h, 2))
        self._peers: list | None = None  # loopback: every rank's handle (rank 0 launches for all of them)
        self._connect()
        self.episode_count, self.episode_sum = 0, 0.0
        self.profile: dict | None = None

    def _connect(self) -> None:
        lib = self._lib
        if isinstance(self.tp, LoopbackTransport):
            handles = self.tp._exchange(int(self._h.value))
            for g, hg in enumerate(handles):
                if g != self.rank:
                    capi.check(lib.qe_shard_connect_local(self._h, g, C.c_void_p(hg)))
            self._peers = handles
            self.tp.barrier()
            return
        if self.world == 1:
            return
        import torch.distributed as dist

        if self._symm is not None:
            ptrs = list(self._symm[1].buffer_ptrs)
            for g in range(self.world):
                if g != self.rank:
                    capi.check(lib.qe_shard_connect_ptr(self._h, g, C.c_void_p(int(ptrs[g]))))
            self.tp.barrier()
            return
        buf = (C.c_ubyte * 64)()
        capi.check(lib.qe_shard_ipc_handle(self._h, buf))
        mine = bytes(buf)
        every = [None] * self.world
        dist.all_gather_object(every, mine, group=getattr(self.tp, "group", None))
        for g, hb in enumerate(every):
            if g != self.rank:
                raw = (C.c_ubyte * 64).from_buffer_copy(hb)
                capi.check(lib.qe_shard_connect_ipc(self._h, g, raw))
        self.tp.barrier()

    def __del__(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.qe_shard_destroy(h)
            except Exception:  # noqa: BLE001
                pass

    def close(self) -> None:
        """Collective: every rank stops using the peers' memory before anybody frees it."""
        _torch().cuda.synchronize()
        self.tp.barrier()
        h, self._h = self._h, None
        if h:
            self._lib.qe_shard_destroy(h)
        self.tp.barrier()
        self._symm = None

    # -- helpers
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream().cuda_stream)

    def fill_random(self, seed: int = 1) -> None:
        """Same values as the single-GPU ``fill_random``: every shard fills its slice of the whole table."""
        capi.check(self._lib.qe_shard_fill_random(self._h, seed, self._stream()))

    def reset(self) -> None:
        """Initial states of this rank's agents (the hash MDP's reset draw by global id)."""
        capi.check(self._lib.qe_shard_reset(self._h, self.seed, T_INIT, self._stream()))
        _torch().cuda.synchronize()
        self.tp.barrier()

    def _launch(self, thresholds: np.ndarray, lrs: np.ndarray) -> None:
        lib, k = self._lib, int(thresholds.shape[0])
        args = (k, thresholds.ctypes.data_as(C.c_void_p), lrs.ctypes.data_as(C.c_void_p), self.seed, self.seed,
                int(self.action_size > 10), 1, self.term_threshold, self._stream())
        if self._peers is not None:  # all ranks on this GPU: one cooperative launch, issued by rank 0
            self.tp.barrier()
            if self.rank == 0:
                arr = (C.c_void_p * self.world)(*[C.c_void_p(hg) for hg in self._peers])
                capi.check(lib.qe_shard_steps(arr, self.world, *args))
                _torch().cuda.synchronize()
            self.tp.barrier()
            capi.check(lib.qe_shard_sync(self._h, self._stream()))
        else:
            arr = (C.c_void_p * 1)(self._h)
            capi.check(lib.qe_shard_steps(arr, 1, *args))

    def step(self, eps: float, lr: float) -> None:
        self._launch(np.asarray([explore_threshold(eps)], dtype=np.uint64), np.asarray([lr], dtype=np.float32))

    def run_steps(self, steps: int, exploration_rate_schedule, lr_schedule) -> None:
        """``steps`` vector steps with the reference's schedule protocol (values read, then ``update(N)``, BRT:245-263);
        ``launch_steps`` of them per kernel launch."""
        done = 0
        while done < steps:
            k = min(self.launch_steps, steps - done)
            th = np.empty(k, dtype=np.uint64)
            lrs = np.empty(k, dtype=np.float32)
            for j in range(k):
                th[j] = explore_threshold(exploration_rate_schedule.get_value())
                lrs[j] = np.float32(lr_schedule.get_value())
                lr_schedule.update(self.num_agents)
                exploration_rate_schedule.update(self.num_agents)
            self._launch(th, lrs)
            done += k

    def phase_us(self, steps: int = 8) -> dict:
        """Per-phase microseconds of the last launch on THIS rank (mean over its first ``steps`` vector steps)."""
        buf = (C.c_uint64 * 128)()
        self.sync()
        capi.check(self._lib.qe_shard_phase_ns(self._h, buf))
        k = max(1, min(int(steps), 16, self.launch_steps))
        t = np.asarray(list(buf), dtype=np.float64).reshape(16, 8)[:k]
        d = lambda a, b: float(np.mean(t[:, a] - t[:, b]) / 1e3)  # noqa: E731
        return {"A_work": d(7, 0), "A_barrier": d(1, 7), "sort_scatter": d(4, 1), "T": d(2, 4), "C": d(3, 2), "sort_local": d(6, 3), "step": d(6, 0)}

    def sync(self) -> None:
        """Wait for this rank's kernels and raise deferred device errors."""
        capi.check(self._lib.qe_shard_sync(self._h, self._stream()))

    def _download(self):
        table = np.empty((max(self.hi - self.lo, 1), self.action_size), dtype=np.float32)
        states = np.empty(max(self.n_here, 1), dtype=np.int32)
        rets = np.empty(max(self.n_here, 1), dtype=np.float32)
        es, ec = C.c_double(0.0), C.c_uint64(0)
        capi.check(self._lib.qe_shard_sync(self._h, self._stream()))
        capi.check(self._lib.qe_shard_download(self._h, table.ctypes.data_as(C.c_void_p), states.ctypes.data_as(C.c_void_p),
                                               rets.ctypes.data_as(C.c_void_p), C.byref(es), C.byref(ec)))
        self.episode_sum, self.episode_count = float(es.value), int(ec.value)
        return table[: self.hi - self.lo], states[: self.n_here], rets[: self.n_here]

    def local_table(self) -> np.ndarray:
        return self._download()[0]

    def rows(self, states) -> np.ndarray:
        """Rows of THIS shard by global state id (parity checks at table sizes that do not fit the host)."""
        st = np.ascontiguousarray(states, dtype=np.int64)
        out = np.empty((st.shape[0], self.action_size), dtype=np.float32)
        capi.check(self._lib.qe_shard_sync(self._h, self._stream()))
        capi.check(self._lib.qe_shard_rows_host(self._h, st.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), st.shape[0]))
        return out

    # -- whole-job views (collective)
    def gather_agents(self):
        """``(states, episode_returns)`` of all agents in global-id order, on every rank (collective)."""
        torch = _torch()
        _, states, rets = self._download()
        rows = torch.stack([torch.from_numpy(states.copy()), torch.from_numpy(rets.copy()).view(torch.int32)], dim=1).to(self.dev)
        allr = self.tp.all_gather_rows(rows)  # rank order == global-id order (contiguous id ranges)
        return allr[:, 0].cpu().numpy(), allr[:, 1].contiguous().view(torch.float32).cpu().numpy()

    def gather_table(self) -> np.ndarray:
        """The whole table ``[S, A]`` on every rank (collective; tests and checkpoints)."""
        torch = _torch()
        part = torch.from_numpy(self._download()[0].copy()).to(self.dev)
        return self.tp.all_gather_rows(part).cpu().numpy()[: self.state_size]

    def table_checksum(self) -> int:
        """Order-free 64-bit checksum of the whole table (sum of the cells' bit patterns times a position hash), reduced
        over the ranks on the device: equal tables <=> equal checksums for all practical purposes (collective)."""
        torch = _torch()
        part = torch.from_numpy(self._download()[0].copy()).to(self.dev)
        idx = torch.arange(self.lo * self.action_size, self.lo * self.action_size + part.numel(), dtype=torch.int64, device=self.dev)
        mix = (idx * 0x1E3779B97F4A7C15 + 0x7F4A7C15) & 0x7FFFFFFFFFFFFFFF
        acc = (part.reshape(-1).view(torch.int32).to(torch.int64) * (mix | 1)).sum().reshape(1)
        self.tp.all_reduce_sum_(acc)
        return int(acc.item()) & 0xFFFFFFFFFFFFFFFF


# ----------------------------------------------------------------------------------------------------- replicated table
class ReplicatedQLearning:
    """Every rank trains its own agents on a full copy of the table; the copies are merged every ``sync_every``
    vector steps by an all-reduce of their deltas (config 5).  ``runtime`` is a single-GPU trainer
    (``SingleThreadQLearning``) whose ``run_steps`` executes the fused loop."""

    def __init__(self, runtime, transport: Transport, sync_every: int = 8, carry_over: bool = False) -> None:
        torch = _torch()
        self.runtime, self.tp, self.sync_every = runtime, transport, int(sync_every)
        algo = runtime.algorithm
        self.algorithm = algo
        self.dev = torch.device("cuda", algo.device)
        shape = (algo.state_size, algo.action_size)
        self.base = torch.empty(shape, dtype=torch.float32, device=self.dev)
        self.delta = torch.empty(shape, dtype=torch.float32, device=self.dev)
        self.syncs = 0
        # carry_over: the period counts vector steps ACROSS run_steps calls (a caller that advances one step per call
        # still merges every sync_every steps) instead of merging at the end of every call
        self.carry_over, self._since_sync = bool(carry_over), 0
        # (Round 1 pinned the per-step-sort form here so that ranks probing different forms would not wait for each other
        # at the all-reduce; the engine's default form, the one-pass pipeline, is one form on every rank.)
        self.rebase()

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream().cuda_stream)

    def rebase(self) -> None:
        """Take the current table as the common base (all ranks must hold the same table at this point)."""
        self.algorithm._before_device_op()
        capi.check(capi.lib().qe_table_export_dense(self.algorithm.handle, self.base.data_ptr(), self._stream()))

    def sync(self) -> None:
        lib, h = capi.lib(), self.algorithm.handle
        self.algorithm._before_device_op()
        trace = getattr(self, "trace_events", None)  # development aid: [(start, after delta, after all-reduce, after merge)]
        ev = [_torch().cuda.Event(enable_timing=True) for _ in range(4)] if trace is not None else None
        if ev:
            ev[0].record()
        capi.check(lib.qe_table_delta_dense(h, self.base.data_ptr(), self.delta.data_ptr(), self._stream()))
        if ev:
            ev[1].record()
        self.tp.all_reduce_sum_(self.delta)
        if ev:
            ev[2].record()
        capi.check(lib.qe_table_merge_dense(h, self.base.data_ptr(), self.delta.data_ptr(), self._stream()))
        if ev:
            ev[3].record()
            trace.append(ev)
        self.algorithm._device_wrote()
        self.syncs += 1

    def run_steps(self, steps: int, env, curr_state_dict=None):
        """Same contract as ``SingleThreadQLearning.run_steps`` (STR:28-76); the tables are merged after every
        ``sync_every`` steps and once more at the end."""
        history: list[float] = []
        state = curr_state_dict
        done = 0
        ep_count, ep_sum = 0, 0.0  # summary mode: episode statistics accumulated over the windows
        while done < steps:
            k = min(self.sync_every - self._since_sync, steps - done)
            _, hist, env, state = self.runtime.run_steps(k, env, state, _mean=False)
            history.extend(hist)
            ep_count += int(getattr(self.runtime, "last_episode_count", 0))
            ep_sum += float(getattr(self.runtime, "last_episode_sum", 0.0))
            done += k
            self._since_sync += k
            if self._since_sync >= self.sync_every or (done == steps and not self.carry_over):
                self.sync()
                self._since_sync = 0
        self.runtime.last_episode_count, self.runtime.last_episode_sum = ep_count, ep_sum
        if getattr(self.runtime, "history_mode", "full") == "full":
            mean = float(sum(history) / len(history))  # the reference's contract for the WHOLE call: ZeroDivisionError if no episode ended (STR:67)
        else:
            mean = ep_sum / ep_count if ep_count else 0.0
        return mean, history, env, state
