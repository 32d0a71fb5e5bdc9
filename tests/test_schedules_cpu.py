"""The schedules (SURVEY a-15) against the UNMODIFIED reference's (baseline/_ref, or /root/reference/src in the build
container): same values, bit for bit, step after step -- as python floats and as shared fp32 cells (``set_mp``, what the
reference's parallel trainer uses)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"):
    if os.path.isdir(os.path.join(cand, "dist_classicrl", "schedules")):
        sys.path.insert(0, cand)
        break
ref_c = pytest.importorskip("dist_classicrl.schedules.constant_schedule")
ref_l = pytest.importorskip("dist_classicrl.schedules.linear_schedule")
ref_e = pytest.importorskip("dist_classicrl.schedules.exponential_schedule")

from dist_classicrl_b200 import schedules as ours  # noqa: E402

CASES = [
    ("ConstantSchedule", ref_c, (0.1,)),
    ("LinearSchedule", ref_l, (1.0, 1.0)),
    ("LinearSchedule", ref_l, (0.5, -1e-4)),
    ("ExponentialSchedule", ref_e, (0.1, 1e-5, 0.995)),   # the benchmark's learning rate (TPB:157-161)
    ("ExponentialSchedule", ref_e, (1.0, 0.01, 0.995)),   # ... and exploration rate
    ("ExponentialSchedule", ref_e, (0.7, 0.05, 0.9999)),
]


@pytest.mark.parametrize("shared", [False, True])
@pytest.mark.parametrize("name,ref_mod,args", CASES)
def test_schedule_values_equal_the_reference(name, ref_mod, args, shared):
    a, b = getattr(ours, name)(*args), getattr(ref_mod, name)(*args)
    if shared:
        a.set_mp()
        b.set_mp()
    assert a.min_value == b.min_value
    for n in [1, 128, 1, 7, 1 << 20, 3, 128, 128, 1, 50_000] * 3:
        assert a.get_value() == b.get_value()
        a.update(n)
        b.update(n)
    assert a.get_value() == b.get_value()


def test_peek_is_k_steps_of_read_then_update():
    a, b = ours.ExponentialSchedule(1.0, 0.01, 0.995), ours.ExponentialSchedule(1.0, 0.01, 0.995)
    seen = a.peek(128, 12)
    expect = []
    for _ in range(12):
        expect.append(b.get_value())
        b.update(128)
    assert seen == expect and a.get_value() == b.get_value()


def test_shared_cells_can_be_adopted_and_import_paths_match_the_reference():
    from dist_classicrl_b200.schedules.base_schedules import BaseSchedule
    from dist_classicrl_b200.schedules.constant_schedule import ConstantSchedule
    from dist_classicrl_b200.schedules.exponential_schedule import ExponentialSchedule
    from dist_classicrl_b200.schedules.linear_schedule import LinearSchedule

    assert issubclass(ConstantSchedule, BaseSchedule) and issubclass(LinearSchedule, BaseSchedule) and issubclass(ExponentialSchedule, BaseSchedule)
    a, b = LinearSchedule(1.0, 0.5), LinearSchedule(9.0, 0.5)
    a.set_mp()
    b.set_mp()
    b.set_value(a.value)      # adopt a's cell (PRT:65-66 hands the cells to the worker processes)
    a.update(2)
    assert b.get_value() == 2.0
    with pytest.raises(AssertionError):
        LinearSchedule(1.0, 0.5).set_value(a.value)   # a plain schedule cannot adopt a shared cell
    with pytest.raises(AssertionError):
        a.set_mp()                                     # already shared
