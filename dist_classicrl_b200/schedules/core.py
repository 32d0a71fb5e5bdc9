"""Host-side schedules of the learning rate and the exploration rate.

The kernels never see a schedule object: the runtimes read one value per vector step (``get_value``), hand it to the
launch (``qe_run_t.explore_thresholds_host`` / ``learning_rates_host``, one entry per fused step -- ``peek`` produces
the K values of a K-step launch) and advance the schedule by the number of table updates of that step (``update(n)``),
which is the reference's protocol (``BaseRuntime._learn``, BRT:245-263).

Arithmetic and surface follow the reference's ``schedules`` package so that a schedule can be swapped between the two
code bases: ``value`` / ``min_value`` attributes, ``set_mp`` (the scalar moves into a process-shared C ``float`` --
which quantises it to fp32, and the reference's parallel trainer relies on exactly that), ``set_value`` (also accepts a
shared cell to adopt), ``get_value``, ``update``.  Rules: constant (constant_schedule.py:6-12), linear
``v + n * rate`` (linear_schedule.py:6-31), exponential ``max(v * rate**n, floor)`` (exponential_schedule.py:6-31).
"""

from __future__ import annotations

from multiprocessing import Value
from multiprocessing.sharedctypes import Synchronized


class BaseSchedule:
    """One scalar and a rule that moves it after ``n`` table updates (``_advance``)."""

    def __init__(self, value: float, min_value: float) -> None:
        self.value: float | Synchronized = value
        self.min_value = min_value

    # -- the scalar: a python float, or a shared fp32 cell after set_mp()
    @property
    def is_shared(self) -> bool:
        return isinstance(self.value, Synchronized)

    def set_mp(self) -> None:
        assert isinstance(self.value, float), "Learning rate must be a float."
        self.value = Value("f", self.value)

    def get_value(self) -> float:
        cell = self.value
        return cell.value if isinstance(cell, Synchronized) else cell

    def set_value(self, value) -> None:
        if isinstance(value, Synchronized):  # adopt somebody else's shared cell
            assert self.is_shared, "self.value must be a multiprocessing Value."
            self.value = value
            return
        if self.is_shared:
            self.value.value = value  # rounds to fp32
        else:
            self.value = value

    # -- the rule
    def _advance(self, current: float, n_updates: int) -> float:
        raise NotImplementedError("This method should be implemented in subclasses.")

    def update(self, steps: int) -> None:
        self.set_value(self._advance(self.get_value(), steps))

    def peek(self, n_updates: int, vector_steps: int) -> list[float]:
        """The values ``vector_steps`` consecutive vector steps of ``n_updates`` agents would read; the schedule ends up
        where those steps would have left it."""
        seen = []
        for _ in range(vector_steps):
            seen.append(self.get_value())
            self.update(n_updates)
        return seen


class ConstantSchedule(BaseSchedule):
    def __init__(self, value: float) -> None:
        super().__init__(value, value)

    def update(self, steps: int) -> None:  # nothing moves (and a shared cell is not rewritten)
        return None

    def _advance(self, current: float, n_updates: int) -> float:
        return current


class LinearSchedule(BaseSchedule):
    """``v <- v + n * decay_rate`` (no floor in the reference: ``min_value`` is a formality)."""

    def __init__(self, value: float, decay_rate: float) -> None:
        super().__init__(value, -1e9)
        self.decay_rate = decay_rate

    def _advance(self, current: float, n_updates: int) -> float:
        return current + n_updates * self.decay_rate


class ExponentialSchedule(BaseSchedule):
    """``v <- max(v * decay_rate**n, min_value)``."""

    def __init__(self, value: float, min_value: float, decay_rate: float) -> None:
        super().__init__(value, min_value)
        self.decay_rate = decay_rate

    def _advance(self, current: float, n_updates: int) -> float:
        return max(current * (self.decay_rate**n_updates), self.min_value)
