"""Learning-rate / exploration schedules (host-side scalars fed to the kernels once per vector step).

Same classes and arithmetic as the reference's ``schedules/`` package (base_schedules.py:8-74,
constant_schedule.py:6-12, linear_schedule.py:6-31, exponential_schedule.py:6-31): values are python floats
(fp64), or a ``multiprocessing.Value("f")`` after ``set_mp()`` -- which quantises them to fp32 exactly like the
reference's parallel runtime does.
"""

from dist_classicrl_b200.schedules.base_schedules import BaseSchedule
from dist_classicrl_b200.schedules.constant_schedule import ConstantSchedule
from dist_classicrl_b200.schedules.exponential_schedule import ExponentialSchedule
from dist_classicrl_b200.schedules.linear_schedule import LinearSchedule

__all__ = ["BaseSchedule", "ConstantSchedule", "ExponentialSchedule", "LinearSchedule"]
