"""Exponential schedule ``v = max(v * decay**n, min)`` (reference: schedules/exponential_schedule.py:6-31)."""

from dist_classicrl_b200.schedules.base_schedules import BaseSchedule


class ExponentialSchedule(BaseSchedule):
    def __init__(self, value: float, min_value: float, decay_rate: float) -> None:
        super().__init__(value, min_value)
        self.decay_rate = decay_rate

    def update(self, steps: int) -> None:
        self.set_value(max(self.get_value() * (self.decay_rate**steps), self.min_value))
