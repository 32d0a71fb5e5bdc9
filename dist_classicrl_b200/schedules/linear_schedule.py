"""Linear schedule ``v += n * rate`` (reference: schedules/linear_schedule.py:6-31)."""

from dist_classicrl_b200.schedules.base_schedules import BaseSchedule


class LinearSchedule(BaseSchedule):
    def __init__(self, value: float, decay_rate: float) -> None:
        super().__init__(value=value, min_value=-1e9)
        self.decay_rate = decay_rate

    def update(self, steps: int) -> None:
        self.set_value(self.get_value() + steps * self.decay_rate)
