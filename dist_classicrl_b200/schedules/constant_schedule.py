"""Import path of the reference (``schedules/constant_schedule.py``); the class lives in ``schedules/core.py``."""

from dist_classicrl_b200.schedules.core import ConstantSchedule

__all__ = ["ConstantSchedule"]
