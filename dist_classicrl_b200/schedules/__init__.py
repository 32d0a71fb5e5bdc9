"""Learning-rate / exploration schedules (host-side scalars fed to the kernels once per vector step); see ``core.py``."""

from dist_classicrl_b200.schedules.core import BaseSchedule, ConstantSchedule, ExponentialSchedule, LinearSchedule

__all__ = ["BaseSchedule", "ConstantSchedule", "ExponentialSchedule", "LinearSchedule"]
