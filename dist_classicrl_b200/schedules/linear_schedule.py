"""Import path of the reference (``schedules/linear_schedule.py``); the class lives in ``schedules/core.py``."""

from dist_classicrl_b200.schedules.core import LinearSchedule

__all__ = ["LinearSchedule"]
