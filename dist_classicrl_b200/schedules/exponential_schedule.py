"""Import path of the reference (``schedules/exponential_schedule.py``); the class lives in ``schedules/core.py``."""

from dist_classicrl_b200.schedules.core import ExponentialSchedule

__all__ = ["ExponentialSchedule"]
