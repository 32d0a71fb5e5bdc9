"""``OptimalQLearningBase`` backed by the B200 engine (drop-in for the reference class of the same name).

Reference: ``src/dist_classicrl/algorithms/base_algorithms/q_learning_optimal.py`` (QLO).  Same constructor,
attributes (``state_size, action_size, discount_factor, q_table, _rng, _np_rng``, QLO:77-98) and methods; the
Q-table lives in HBM as fp32 (the reference defaults to fp64 -- SURVEY 0.4 -- every formula is evaluated in
fp32 with one rounding per operation, exactly what the reference does with a float32 table and float32
rewards).  ``choose_actions`` / ``learn`` accept host NumPy arrays (copied through the C ABI's ``*_host`` entry
points) or CUDA ``torch`` tensors (zero-copy).

Randomness.  The reference draws from ``self._rng`` (``random.Random``) / ``self._np_rng``.  Here ``_rng`` is a
:class:`~dist_classicrl_b200.rng.CounterRNG` by default: a seed and a step counter; the kernels evaluate the
uniform stream ``U[t, i, k]`` themselves.  Assigning a :class:`~dist_classicrl_b200.rng.PredrawnUniforms` feeds
caller-supplied numbers.  Any other object assigned to ``_rng`` (the reference's tests inject duck-typed shims,
T-RT:17-45, T-QLO ``patch.object(ql._rng, ...)``) is honoured call by call, in the reference's call order: the
table rows are gathered on the GPU and the RNG protocol of the dispatcher branch (QLO:644-726) runs on the host.
"""

from __future__ import annotations

import ctypes as C
import math
from typing import TYPE_CHECKING

import numpy as np

from dist_classicrl_b200 import capi
from dist_classicrl_b200.rng import CounterRNG, PredrawnUniforms, explore_threshold, is_engine_rng

if TYPE_CHECKING:
    from numpy.typing import NDArray

# dispatcher thresholds of the reference (QLO:14-20); they only decide which RNG methods are called and
# how an all-zero action mask behaves
NUM_STATES_LEARN_THRESHOLD = 10
DETERMINISTIC_MAX_ACTION_SIZE_ITER = 10
DETERMINISTIC_MIN_ACTION_SIZE_VEC_ITER = 10000
DETERMINISTIC_MAX_NUM_STATES_VEC_ITER = 3
NO_ACTION_MASKS_NO_DETERMINISTIC_MAX_NUM_STATES_ITER = 100
NO_ACTION_MASKS_NO_DETERMINISTIC_MIN_ACTION_SIZE_VEC_ITER = 100
ACTION_MASKS_NO_DETERMINISTIC_MAX_ACTION_SIZE_ITER = 10

_ITER, _VEC_ITER, _VEC = "iter", "vec_iter", "vec"


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())


def masks_to_bits(masks, action_size: int) -> NDArray[np.uint32]:
    """``[N, A]`` truthy array -> ``uint32[N]`` (bit a = action a legal); A <= 32."""
    m = np.asarray(masks) != 0
    assert m.ndim == 2 and m.shape[1] == action_size, "Action masks must match the number of states and actions."
    packed = np.packbits(m, axis=1, bitorder="little")
    out = np.zeros((m.shape[0], 4), dtype=np.uint8)
    out[:, : packed.shape[1]] = packed
    return out.view("<u4").reshape(-1)


class OptimalQLearningBase:
    """Tabular Q-learning with the table resident on one B200 (see module docstring)."""

    state_size: int
    action_size: int
    discount_factor: float

    def __init__(
        self,
        state_size: int | np.integer,
        action_size: int | np.integer,
        discount_factor: float = 0.97,
        seed: int | None = None,
        device: int | None = None,
    ) -> None:
        self.state_size = int(state_size)
        self.action_size = int(action_size)
        self.discount_factor = discount_factor
        self._lib = capi.lib()
        if device is None:
            try:
                import torch

                device = torch.cuda.current_device() if torch.cuda.is_available() else 0
            except ImportError:  # torch is plumbing, not a requirement of the single-GPU path
                device = 0
        self.device = int(device)
        self._h = C.c_void_p()
        capi.check(self._lib.qe_create(self.state_size, self.action_size, float(discount_factor), self.device, C.byref(self._h)))
        self._gamma_on_device = float(discount_factor)
        self._np_rng = np.random.default_rng(seed)
        self._rng = CounterRNG(seed)
        self._host: np.ndarray | None = None  # host mirror of the table (lazy)
        self._host_valid = False  # mirror holds the device contents
        self._host_dirty = False  # mirror may have been modified through the q_table property

    def __del__(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.qe_destroy(h)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass

    # ------------------------------------------------------------------ table ownership (SURVEY 8b "Ownership")
    @property
    def q_table(self) -> NDArray[np.float32]:
        """Host view of the table.  Reading it downloads the device table if it changed; the returned array
        may be modified in place -- it is uploaded again before the next device operation."""
        if self._host is None:
            self._host = np.zeros((self.state_size, self.action_size), dtype=np.float32)
        self._refresh_host()
        self._host_dirty = True  # the caller may write through the returned array
        return self._host

    def _refresh_host(self) -> NDArray[np.float32]:
        """The mirror with the device contents, WITHOUT marking it dirty (read-only accessors, ``save``)."""
        if self._host is None:
            self._host = np.zeros((self.state_size, self.action_size), dtype=np.float32)
        if not self._host_valid:
            capi.check(self._lib.qe_table_download_host(self._h, _ptr(self._host)))
            self._host_valid = True
        return self._host

    @q_table.setter
    def q_table(self, value) -> None:
        arr = np.asarray(value)
        if arr.shape != (self.state_size, self.action_size):
            raise ValueError(f"q_table must have shape {(self.state_size, self.action_size)}, got {arr.shape}")
        if arr.dtype != np.float32 or not arr.flags.c_contiguous or not arr.flags.writeable:
            arr = np.ascontiguousarray(arr, dtype=np.float32).copy()
        self._host = arr  # adopted: e.g. the shared-memory array of the parallel runtime (PRT:73-77)
        self._host_valid = True
        self._host_dirty = True

    def _before_device_op(self) -> None:
        if self._host_dirty:
            capi.check(self._lib.qe_table_upload_host(self._h, _ptr(self._host)))
            self._host_dirty = False
        if float(self.discount_factor) != self._gamma_on_device:
            capi.check(self._lib.qe_set_discount(self._h, float(self.discount_factor)))
            self._gamma_on_device = float(self.discount_factor)

    def _device_wrote(self) -> None:
        self._host_valid = False

    @property
    def handle(self) -> C.c_void_p:
        """The ``qe_engine_t*`` (for the fused runtime and the environments)."""
        return self._h

    def table_device_ptr(self) -> tuple[int, int]:
        """(device pointer, row stride in floats) of the fp32 table."""
        self._before_device_op()
        return int(self._lib.qe_table_ptr(self._h)), int(self._lib.qe_table_stride(self._h))

    def fill_random(self, seed: int = 1) -> None:
        """Uniform [0,1) initial table generated on the device (throughput runs)."""
        capi.check(self._lib.qe_table_fill_random(self._h, seed, None))
        self._host_dirty = False
        self._device_wrote()

    # ------------------------------------------------------------------ accessors (QLO:100-261)
    # (read-only accessors go through _refresh_host: they return copies / scalars and do not force a re-upload of the table;
    #  the reference returns views of its array for slices -- callers that write through them must use q_table)
    def get_q_value(self, state: int, action: int) -> float:
        return self._refresh_host()[state, action]

    def get_q_values(self, states, actions):
        return self._refresh_host()[states, actions]

    def get_state_q_values(self, state: int):
        return self._refresh_host()[state].copy()

    def get_states_q_values(self, states):
        return self._refresh_host()[states]

    def get_action_q_values(self, action: int):
        return self._refresh_host()[:, action].copy()

    def get_actions_q_values(self, actions):
        return self._refresh_host()[:, actions]

    def set_q_value(self, state: int, action: int, value: float) -> None:
        self.q_table[state, action] = value

    def add_q_value(self, state: int, action: int, value: float) -> None:
        self.q_table[state, action] += value

    def add_q_values(self, states, actions, values) -> None:
        np.add.at(self.q_table, (states, actions), values)  # duplicates accumulate (T-QLO:64-82)

    def save(self, filename: str) -> None:
        np.save(filename, self._refresh_host())

    def load(self, filename: str) -> None:
        """Resume from a table written by :meth:`save` (or by the reference's ``save``, QLO:252-261); an fp64 table
        of the reference is rounded to fp32."""
        name = filename if str(filename).endswith(".npy") else f"{filename}.npy"
        self.q_table = np.load(name)

    # ------------------------------------------------------------------ select
    def _variant(self, n: int, deterministic: bool, masked: bool) -> str:
        """Dispatcher branch of QLO:671-726 (decides RNG method names and the empty-mask behaviour)."""
        a = self.action_size
        if deterministic:
            if a <= DETERMINISTIC_MAX_ACTION_SIZE_ITER:
                return _ITER
            if a >= DETERMINISTIC_MIN_ACTION_SIZE_VEC_ITER and n <= DETERMINISTIC_MAX_NUM_STATES_VEC_ITER:
                return _VEC_ITER
            return _VEC
        if not masked:
            if n < NO_ACTION_MASKS_NO_DETERMINISTIC_MAX_NUM_STATES_ITER:
                return _ITER
            if a > NO_ACTION_MASKS_NO_DETERMINISTIC_MIN_ACTION_SIZE_VEC_ITER:
                return _VEC_ITER
            return _VEC
        return _ITER if a <= ACTION_MASKS_NO_DETERMINISTIC_MAX_ACTION_SIZE_ITER else _VEC_ITER

    def choose_actions(
        self,
        states,
        exploration_rate: float,
        *,
        deterministic: bool = False,
        action_masks=None,
    ):
        """Masked epsilon-greedy actions for all agents (QLO:644-726).  Returns ``int32[N]`` (``-1`` where an
        agent has no candidate in the iter variants, QLO:348)."""
        n = len(states)
        return self._choose(states, exploration_rate, deterministic, action_masks, self._variant(n, deterministic, action_masks is not None))

    @staticmethod
    def _checked(idx, size: int, what: str) -> np.ndarray:
        """Host index array as int32 inside [0, size): negative values wrap NumPy-style, anything else out of range raises
        IndexError like the reference's ``q_table[idx]`` would -- the kernels index HBM directly and do not check."""
        arr = np.asarray(idx)
        if arr.dtype.kind not in "iu":
            arr = arr.astype(np.int64)
        if arr.size and (arr.min() < 0 or arr.max() >= size):
            arr = arr.astype(np.int64, copy=True)
            arr[arr < 0] += size
            if arr.size and (arr.min() < 0 or arr.max() >= size):
                bad = arr[(arr < 0) | (arr >= size)][0]
                raise IndexError(f"{what} index {int(bad)} is out of bounds for size {size}")
        return np.ascontiguousarray(arr, dtype=np.int32)

    def _choose(self, states, eps, deterministic, action_masks, variant):
        self._before_device_op()
        n = len(states)
        if n == 0:
            return np.zeros(0, dtype=np.int32)
        empty_all = int(variant != _ITER)
        if not is_engine_rng(self._rng) or type(self._np_rng) is not np.random.Generator:
            # someone injected / patched a generator (T-RT:70, T-MPI:36-37, T-QLO patch.object): honour it
            return self._choose_with_injected_rng(states, eps, deterministic, action_masks, variant)
        rng = self._rng
        t = rng.next_step()
        thresh = explore_threshold(eps)
        a = self.action_size
        if _is_torch_cuda(states):
            import torch

            st = states if states.dtype == torch.int32 else states.to(torch.int32)
            st = st.contiguous()
            bits = mbytes = None
            if action_masks is not None:
                if action_masks.dim() == 1:  # already a uint32 bitmask carried in int32
                    bits = action_masks.contiguous()
                else:
                    mbytes = (action_masks != 0).to(torch.uint8).contiguous()
            out = torch.empty(n, dtype=torch.int32, device=states.device)
            u_dev, slots, seed = None, 2, 0
            if isinstance(rng, PredrawnUniforms):
                u_host = rng.row(t)
                u_dev = torch.from_numpy(u_host.view(np.int32)).to(states.device)
                slots = u_host.shape[1]
            else:
                seed = rng.seed
            stream = torch.cuda.current_stream().cuda_stream
            capi.check(self._lib.qe_select(self._h, _ptr(st), _ptr(bits), _ptr(mbytes), _ptr(u_dev), slots, seed, t, 0, thresh,
                                           int(deterministic), empty_all, _ptr(out), n, C.c_void_p(stream)))
            return out
        st = self._checked(states, self.state_size, "state")
        bits = mbytes = None
        if action_masks is not None:
            if a <= 32:
                bits = masks_to_bits(action_masks, a)
            else:
                mbytes = np.ascontiguousarray(np.asarray(action_masks) != 0, dtype=np.uint8)
                assert mbytes.shape == (n, a), "Action masks must match the number of states and actions."
        out = np.empty(n, dtype=np.int32)
        u, slots, seed = None, 2, 0
        if isinstance(rng, PredrawnUniforms):
            u = np.ascontiguousarray(rng.row(t)[:n])
            slots = u.shape[1]
        else:
            seed = rng.seed
        capi.check(self._lib.qe_select_host(self._h, _ptr(st), _ptr(bits), _ptr(mbytes), _ptr(u), slots, seed, t, 0, thresh,
                                            int(deterministic), empty_all, _ptr(out), n))
        if empty_all and action_masks is not None and (out < 0).any():
            # vec variants call _rng.choice on an empty array when exploring with an all-zero mask (QLO:465,470)
            raise IndexError("Cannot choose from an empty sequence")
        return out

    def _rows(self, states) -> np.ndarray:
        st = self._checked(states, self.state_size, "state")
        rows = np.empty((st.shape[0], self.action_size), dtype=np.float32)
        capi.check(self._lib.qe_gather_rows_host(self._h, _ptr(st), _ptr(rows), st.shape[0]))
        return rows

    def _choose_with_injected_rng(self, states, eps, deterministic, action_masks, variant):
        """Generic path: table rows from the GPU, RNG protocol of the reference on the host, call by call."""
        if _is_torch_cuda(states):
            states = states.cpu().numpy()
            action_masks = None if action_masks is None else action_masks.cpu().numpy()
        rows = self._rows(states)
        n, a = rows.shape
        rng = self._rng
        masks = None if action_masks is None else np.asarray(action_masks)
        if masks is not None:
            assert masks.shape == (n, a), "Action mask should have the same length as the action size."
        out = np.empty(n, dtype=np.int32)
        if variant == _VEC:  # batch variants: QLO:524-579 (no masks), 581-642 (masks)
            if masks is None:
                best = rows.max(axis=1, keepdims=True)
                if deterministic:
                    for i in range(n):
                        out[i] = rng.choice(np.where(rows[i] == best[i])[0])
                    return out
                explore = self._np_rng.random(n) < eps
                exploratory = self._np_rng.integers(a, size=n)
                for i in range(n):
                    out[i] = exploratory[i] if explore[i] else rng.choice(np.where(rows[i] == best[i])[0])
                return out
            masked = np.where(masks, rows, -np.inf)
            best = masked.max(axis=1, keepdims=True)
            explore = np.zeros(n, dtype=bool) if deterministic else self._np_rng.random(n) < eps
            for i in range(n):
                cand = np.where(masks[i])[0] if explore[i] else np.where(masked[i] == best[i])[0]
                out[i] = rng.choice(cand)
            return out
        draw = (lambda: rng.uniform(0, 1)) if variant == _ITER else rng.random  # QLO:287,335 vs QLO:426,464
        for i in range(n):
            explore = (not deterministic) and draw() < eps
            row = rows[i]
            if masks is None:
                if explore:
                    out[i] = rng.randint(0, a - 1)  # QLO:288, 427
                    continue
                if variant == _ITER:  # QLO:289-302
                    cand = [int(k) for k in np.nonzero(row == row.max())[0]] if not math.isinf(-row.max()) else []
                    out[i] = rng.choice(cand) if cand else -1
                else:  # QLO:429-430
                    out[i] = rng.choice(np.where(row == row.max())[0])
                continue
            mask = masks[i]
            if variant == _ITER:  # QLO:331-348
                if explore:
                    cand = [k for k in range(a) if mask[k]]
                else:
                    legal = [k for k in range(a) if mask[k]]
                    top = max((row[k] for k in legal), default=-math.inf)
                    cand = [k for k in legal if row[k] == top and top > -math.inf]
                out[i] = rng.choice(cand) if cand else -1
            else:  # QLO:459-470
                np_mask = np.fromiter(mask, dtype=np.int32, count=len(mask))
                if explore:
                    cand = np.where(np_mask)[0]
                else:
                    masked_row = np.where(np_mask, row, -np.inf)
                    cand = np.where(masked_row == masked_row.max())[0]
                out[i] = rng.choice(cand)
        return out

    # the reference's public variants -- all evaluate the same function
    def choose_actions_iter(self, states, exploration_rate, *, deterministic=False, action_masks=None):
        return self._choose(states, exploration_rate, deterministic, action_masks, _ITER)

    def choose_actions_vec_iter(self, states, exploration_rate, *, deterministic=False, action_masks=None):
        return self._choose(states, exploration_rate, deterministic, action_masks, _VEC_ITER)

    def choose_actions_vec(self, states, exploration_rate, *, deterministic=False):
        return self._choose(states, exploration_rate, deterministic, None, _VEC)

    def choose_masked_actions_vec(self, states, action_masks, exploration_rate, *, deterministic=False):
        assert np.asarray(action_masks).shape == (len(states), self.action_size), (
            "Action masks must match the number of states and actions."
        )
        return self._choose(states, exploration_rate, deterministic, action_masks, _VEC)

    def choose_action(self, state: int, exploration_rate: float, *, deterministic: bool = False) -> int:
        return int(self._choose(np.asarray([state]), exploration_rate, deterministic, None, _ITER)[0])

    def choose_masked_action(self, state: int, action_mask, exploration_rate: float, *, deterministic: bool = False) -> int:
        assert len(action_mask) == self.action_size, "Action mask should have the same length as the action size."
        return int(self._choose(np.asarray([state]), exploration_rate, deterministic, np.asarray([action_mask]), _ITER)[0])

    def choose_action_vec(self, state: int, exploration_rate: float, *, deterministic: bool = False) -> int:
        return int(self._choose(np.asarray([state]), exploration_rate, deterministic, None, _VEC_ITER)[0])

    def choose_masked_action_vec(self, state: int, action_mask, exploration_rate: float, *, deterministic: bool = False) -> int:
        assert len(action_mask) == self.action_size, "Action mask should have the same size as the action space."
        return int(self._choose(np.asarray([state]), exploration_rate, deterministic, np.asarray([action_mask]), _VEC_ITER)[0])

    # ------------------------------------------------------------------ learn
    def _learn(self, states, actions, rewards, next_states, terminated, lr, next_action_masks, mode) -> None:
        self._before_device_op()
        n = len(states)
        if n == 0:
            return
        a = self.action_size
        lr32 = float(np.float32(lr))
        if _is_torch_cuda(states):
            import torch

            def as_t(x, dt):
                x = x if x.dtype == dt else x.to(dt)
                return x.contiguous()

            bits = mbytes = None
            if next_action_masks is not None:
                if next_action_masks.dim() == 1:
                    bits = next_action_masks.contiguous()
                else:
                    mbytes = (next_action_masks != 0).to(torch.uint8).contiguous()
            args = (as_t(states, torch.int32), as_t(actions, torch.int32), as_t(rewards, torch.float32),
                    as_t(next_states, torch.int32), as_t(terminated, torch.uint8))
            stream = torch.cuda.current_stream().cuda_stream
            capi.check(self._lib.qe_learn(self._h, *[_ptr(x) for x in args], _ptr(bits), _ptr(mbytes), lr32, n, mode,
                                          C.c_void_p(stream)))
            self._device_wrote()
            return
        bits = mbytes = None
        if next_action_masks is not None:
            if a <= 32:
                bits = masks_to_bits(next_action_masks, a)
            else:
                mbytes = np.ascontiguousarray(np.asarray(next_action_masks) != 0, dtype=np.uint8)
        args = (self._checked(states, self.state_size, "state"), self._checked(actions, a, "action"),
                np.ascontiguousarray(rewards, dtype=np.float32), self._checked(next_states, self.state_size, "next state"),
                np.ascontiguousarray(terminated, dtype=np.uint8))
        for arr in args[1:]:
            if arr.shape[0] != n:  # zip(strict=True) in QLO:801-808
                raise ValueError("learn(): argument lengths differ")
        self._device_wrote()
        capi.check(self._lib.qe_learn_host(self._h, *[_ptr(x) for x in args], _ptr(bits), _ptr(mbytes), lr32, n, mode))

    def learn(self, states, actions, rewards, next_states, terminated, lr: float, next_action_masks=None) -> None:
        """Sequential per-agent TD update, agent *i* sees the writes of agents *j < i* (QLO:893-934)."""
        self._learn(states, actions, rewards, next_states, terminated, lr, next_action_masks, capi.QE_LEARN_SEQUENTIAL)

    def learn_iter(self, states, actions, rewards, next_states, terminated, lr: float, next_action_masks=None) -> None:
        self._learn(states, actions, rewards, next_states, terminated, lr, next_action_masks, capi.QE_LEARN_SEQUENTIAL)

    def single_learn(self, state, action, reward, next_state, terminated, lr, next_action_mask=None) -> None:
        masks = None if next_action_mask is None else np.asarray([next_action_mask])
        self._learn(np.asarray([state]), np.asarray([action]), np.asarray([reward], dtype=np.float32),
                    np.asarray([next_state]), np.asarray([bool(terminated)]), lr, masks, capi.QE_LEARN_SEQUENTIAL)

    def learn_vec(self, states, actions, rewards, next_states, terminated, lr: float, next_action_masks=None) -> None:
        """Snapshot bootstrap + accumulating scatter (``np.add.at`` semantics, QLO:853-891)."""
        self._learn(states, actions, rewards, next_states, terminated, lr, next_action_masks, capi.QE_LEARN_ACCUMULATE)

    _learn_vec = learn_vec
