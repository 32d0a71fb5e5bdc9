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
    """State-range-sharded Q-table with the hash MDP's agents living on the shard of their state (config 4).

    Parameters mirror ``OptimalQLearningBase`` (QLO:84-90) plus the environment of ``HashMDPVecEnv``.  Results
    (table, agent states, episode returns) are identical to the single-GPU engine run on the same seeds.
    """

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
        capi.check(self._lib.qe_create(max(self.hi - self.lo, 1), self.action_size, self.discount_factor, self.device_index, C.byref(self._h)))
        capi.check(self._lib.qe_set_state_base(self._h, self.lo))
        self.gamma32 = torch.tensor(np.float32(discount_factor), dtype=torch.float32, device=self.dev)
        self.t = 0  # vector-step counter of both uniform streams
        self.gid = self.state = self.ep_ret = None
        self.episode_count, self.episode_sum = 0, 0.0
        self.rounds_last = 0
        self.rounds_total = 0
        self.profile: dict | None = None  # set to {} to collect per-section wall times (synchronising; diagnostics only)

    def __del__(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.qe_destroy(h)
            except Exception:  # noqa: BLE001
                pass

    # -- helpers
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream().cuda_stream)

    def fill_random(self, seed: int = 1) -> None:
        """Same values as the single-GPU ``fill_random``: every shard fills its slice of the whole table."""
        capi.check(self._lib.qe_table_fill_random(self._h, seed, self._stream()))

    def local_table(self) -> np.ndarray:
        out = np.empty((max(self.hi - self.lo, 1), self.action_size), dtype=np.float32)
        capi.check(self._lib.qe_table_download_host(self._h, out.ctypes.data_as(C.c_void_p)))
        return out[: self.hi - self.lo]

    def _masks(self, states):
        torch = _torch()
        out = torch.empty(states.shape[0], dtype=torch.int32, device=self.dev)
        capi.check(self._lib.qe_mdp_masks(states.data_ptr(), out.data_ptr(), self.action_size, self.env_seed, states.shape[0], self._stream()))
        return out

    def _migrate(self, gid, state, ep_ret):
        torch = _torch()
        rows = torch.stack([gid, state, ep_ret.view(torch.int32)], dim=1)
        recv, _, _ = route(self.tp, rows, owner_of(state, self.state_size, self.world).to(torch.int64))
        order = torch.sort(recv[:, 0], stable=True).indices  # global-id order == the sequential order of the update
        recv = recv[order]
        self.gid = recv[:, 0].contiguous()
        self.state = recv[:, 1].contiguous()
        self.ep_ret = recv[:, 2].contiguous().view(torch.float32)

    # -- environment
    def reset(self) -> None:
        """Initial states of the agents this rank draws (a contiguous global-id range), then migration to the owners."""
        torch = _torch()
        n, g = self.num_agents, self.world
        per = -(-n // g)
        a0, a1 = min(self.rank * per, n), min((self.rank + 1) * per, n)
        cnt = a1 - a0
        state = torch.empty(max(cnt, 1), dtype=torch.int32, device=self.dev)[:cnt]
        if cnt:
            capi.check(self._lib.qe_mdp_reset(state.data_ptr(), None, self.state_size, self.action_size, self.env_seed, None, 4,
                                              self.seed, T_INIT, a0, cnt, self._stream()))
        gid = torch.arange(a0, a1, dtype=torch.int32, device=self.dev)
        self._migrate(gid, state, torch.zeros(cnt, dtype=torch.float32, device=self.dev))
        self.t = 0

    def _mark(self, name: str) -> None:
        if self.profile is not None:
            import time

            _torch().cuda.synchronize()
            now = time.perf_counter()
            self.profile[name] = self.profile.get(name, 0.0) + (now - self._t_mark)
            self._t_mark = now

    # -- one vector step
    def step(self, eps: float, lr: float) -> None:
        torch, lib, h = _torch(), self._lib, self._h
        st = self._stream()
        if self.profile is not None:
            import time

            torch.cuda.synchronize()
            self._t_mark = time.perf_counter()
        n = int(self.gid.shape[0])
        thresh = explore_threshold(eps)
        gid, s_old = self.gid, self.state
        s2 = s_old.clone()
        actions = torch.empty(max(n, 1), dtype=torch.int32, device=self.dev)[:n]
        rewards = torch.empty(max(n, 1), dtype=torch.float32, device=self.dev)[:n]
        term = torch.empty(max(n, 1), dtype=torch.uint8, device=self.dev)[:n]
        mask2 = torch.empty(max(n, 1), dtype=torch.int32, device=self.dev)[:n]
        if n:
            capi.check(lib.qe_set_agent_ids(h, gid.data_ptr()))
            masks = self._masks(s_old)
            capi.check(lib.qe_select(h, s_old.data_ptr(), masks.data_ptr(), None, None, 4, self.seed, self.t, 0, thresh, 0,
                                     int(self.action_size > 10), actions.data_ptr(), n, st))
            capi.check(lib.qe_mdp_step(h, s2.data_ptr(), actions.data_ptr(), self.state_size, self.action_size, self.env_seed,
                                       self.term_threshold, None, 4, self.seed, self.t, 0, mask2.data_ptr(), rewards.data_ptr(),
                                       term.data_ptr(), n, st))
            capi.check(lib.qe_set_agent_ids(h, None))
        self._mark("select+env")
        # episode bookkeeping (BRT:212-221)
        acc = self.ep_ret + rewards
        done = term != 0
        if n:
            self.episode_count += int(done.sum().item())
            self.episode_sum += float(acc[done].double().sum().item())
        self.ep_ret = torch.where(done, torch.zeros_like(acc), acc)

        self._mark("bookkeeping")
        # ---- TD update, exact in global agent order
        own2 = owner_of(s2, self.state_size, self.world).to(torch.int64)
        remote = (~done) & (own2 != self.rank)
        ridx = torch.nonzero(remote).reshape(-1)
        req_rows = torch.stack([gid[ridx], s2[ridx]], dim=1) if n else torch.empty((0, 2), dtype=torch.int32, device=self.dev)
        order, out_counts = bucket_by_owner(own2[ridx], self.world)
        recv = self.tp.all_to_all_v(list(torch.split(req_rows[order], out_counts, dim=0)))
        in_counts = [int(x.shape[0]) for x in recv]
        got = torch.cat(recv, dim=0)
        nreq = int(got.shape[0])
        req_s2 = got[:, 1].contiguous()
        req_pos = torch.searchsorted(gid, got[:, 0].contiguous()).to(torch.int32) if nreq else got[:, 0].contiguous()
        req_mask = self._masks(req_s2) if nreq else req_s2
        answers = torch.empty((max(nreq, 1), 1), dtype=torch.float32, device=self.dev)[:nreq]

        def serve(use_versions: int):
            if nreq:
                capi.check(lib.qe_serve_bootstrap(h, req_s2.data_ptr(), req_pos.data_ptr(), req_mask.data_ptr(), answers.data_ptr(),
                                                  nreq, use_versions, st))
            return route_back(self.tp, answers.view(torch.int32), in_counts, order, out_counts).view(torch.float32).reshape(-1)

        self._mark("route requests")
        m_ext = serve(0)  # snapshot values to start from
        self._mark("serve0")
        term_eff = torch.where(remote, torch.ones_like(term), term)
        s2_safe = torch.where(remote, s_old, s2)  # remote rows are never touched locally
        lr32 = float(np.float32(lr))
        capi.check(lib.qe_set_hold(h, 1))
        rounds = 0
        while True:
            r_eff = rewards.clone()
            if ridx.numel():
                r_eff[ridx] = rewards[ridx] + self.gamma32 * m_ext  # target of a remote bootstrap: r + gamma*m, one rounding each
            if n:
                capi.check(lib.qe_learn(h, s_old.data_ptr(), actions.data_ptr(), r_eff.data_ptr(), s2_safe.data_ptr(), term_eff.data_ptr(),
                                        mask2.data_ptr(), None, lr32, n, capi.QE_LEARN_SEQUENTIAL, st))
            rounds += 1
            self._mark("learn (hold)")
            m_new = serve(1)
            self._mark("serve")
            changed = (m_new.view(torch.int32) != m_ext.view(torch.int32)).any().to(torch.int32).reshape(1)
            m_ext = m_new
            stop = self.tp.all_reduce_max_flag(changed) == 0
            self._mark("converged?")
            if stop:
                break
            if rounds > 4096:
                raise capi.EngineError("sharded TD update did not reach its fixed point")
        if n:
            capi.check(lib.qe_learn_commit(h, s_old.data_ptr(), actions.data_ptr(), n, st))
        capi.check(lib.qe_set_hold(h, 0))
        capi.check(lib.qe_sync(h, st))
        self.rounds_last = rounds
        self.rounds_total += rounds
        self.t = (self.t + 1) & 0xFFFFFFFF
        self._mark("commit")
        # ---- migration to owner(s')
        self._migrate(gid, s2, self.ep_ret)
        self._mark("migrate")

    def run_steps(self, steps: int, exploration_rate_schedule, lr_schedule) -> None:
        """``steps`` vector steps with the reference's schedule protocol (values read, then ``update(N)``, BRT:245-263)."""
        for _ in range(steps):
            eps, lr = exploration_rate_schedule.get_value(), lr_schedule.get_value()
            self.step(eps, lr)
            lr_schedule.update(self.num_agents)
            exploration_rate_schedule.update(self.num_agents)

    # -- whole-job views (collective)
    def gather_agents(self):
        """``(states, episode_returns)`` of all agents in global-id order, on every rank (collective)."""
        torch = _torch()
        rows = torch.stack([self.gid, self.state, self.ep_ret.view(torch.int32)], dim=1)
        allr = self.tp.all_gather_rows(rows)
        allr = allr[torch.sort(allr[:, 0]).indices]
        return allr[:, 1].cpu().numpy(), allr[:, 2].contiguous().view(torch.float32).cpu().numpy()

    def gather_table(self) -> np.ndarray:
        """The whole table ``[S, A]`` on every rank (collective; tests and checkpoints)."""
        torch = _torch()
        part = torch.from_numpy(self.local_table()).to(self.dev)
        return self.tp.all_gather_rows(part).cpu().numpy()[: self.state_size]


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
        # Merged replicas learn world_size times as fast, so agents herd onto the same rows sooner: the per-step sort is
        # the form of the exact update whose cost does not grow with the crowd, and a form that is the same on every rank
        # keeps the ranks in step at the all-reduce (QE_SORTED in the environment overrides).
        if transport.world_size > 1 and "QE_SORTED" not in os.environ:
            capi.check(capi.lib().qe_set_fused_form(algo.handle, 1))
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
        while done < steps:
            k = min(self.sync_every - self._since_sync, steps - done)
            _, hist, env, state = self.runtime.run_steps(k, env, state)
            history.extend(hist)
            done += k
            self._since_sync += k
            if self._since_sync >= self.sync_every or (done == steps and not self.carry_over):
                self.sync()
                self._since_sync = 0
        mean = float(np.mean(history)) if history else 0.0
        return mean, history, env, state
