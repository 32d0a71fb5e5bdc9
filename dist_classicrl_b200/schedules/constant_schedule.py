"""Constant schedule (reference: schedules/constant_schedule.py:6-12)."""

from dist_classicrl_b200.schedules.base_schedules import BaseSchedule


class ConstantSchedule(BaseSchedule):
    def __init__(self, value: float) -> None:
        super().__init__(value=value, min_value=value)

    def update(self, steps: int) -> None:
        return None
