"""Schedule base class (reference: schedules/base_schedules.py:8-74)."""

from __future__ import annotations

from multiprocessing import Value
from multiprocessing.sharedctypes import Synchronized


class BaseSchedule:
    """Holds one scalar; subclasses define how ``update(n)`` moves it after ``n`` table updates."""

    def __init__(self, value: float, min_value: float) -> None:
        self.value = value
        self.min_value = min_value

    def _shared(self) -> bool:
        return isinstance(self.value, Synchronized)

    def set_mp(self) -> None:
        """Move the scalar into process-shared memory (a C ``float``: fp32, base_schedules.py:27-30)."""
        assert isinstance(self.value, float), "Learning rate must be a float."
        self.value = Value("f", self.value)

    def set_value(self, value) -> None:
        if isinstance(value, Synchronized):
            assert self._shared(), "self.value must be a multiprocessing Value."
            self.value = value
        elif self._shared():
            self.value.value = value
        else:
            self.value = value

    def get_value(self) -> float:
        return self.value.value if self._shared() else self.value

    def update(self, steps: int) -> None:
        raise NotImplementedError("This method should be implemented in subclasses.")

    def peek(self, n_updates: int, vector_steps: int) -> list[float]:
        """Values seen by ``vector_steps`` consecutive vector steps of ``n_updates`` agents each; the schedule
        itself is advanced past them (what K iterations of BRT:262-263 would do)."""
        out = []
        for _ in range(vector_steps):
            out.append(self.get_value())
            self.update(n_updates)
        return out
