"""Import path of the reference (``schedules/base_schedules.py``); the class lives in ``schedules/core.py``."""

from dist_classicrl_b200.schedules.core import BaseSchedule

__all__ = ["BaseSchedule"]
