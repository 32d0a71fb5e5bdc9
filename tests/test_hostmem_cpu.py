"""Host-side guards of the in-place page-locking helper (no GPU needed: these cases never reach the driver) and the
batch slicer of the actor/learner trainer."""
import numpy as np


def test_pin_refuses_what_cannot_be_registered():
    from dist_classicrl_b200 import hostmem

    assert hostmem.pin([1.0, 2.0]) is False                                   # not an array
    assert hostmem.pin(np.zeros((8, 8), dtype=np.float32)[:, ::2]) is False   # not contiguous
    assert hostmem.pin(np.zeros(0, dtype=np.float32)) is False                # empty
    ro = np.zeros(1024, dtype=np.float32)
    ro.flags.writeable = False
    assert hostmem.pin(ro) is False                                           # read-only buffers are left alone
    assert hostmem._registered == {} or all(isinstance(k, int) for k in hostmem._registered)


def test_take_slices_queued_records_in_order():
    from dist_classicrl_b200.algorithms.runtime.q_learning_async_dist import _take

    def rec(lo, n, masks=True):
        r = np.arange(lo, lo + n)
        return (r.astype(np.int32), (r % 3).astype(np.int32), r.astype(np.float32), (r + 1).astype(np.int32),
                np.stack([r, r], axis=1).astype(np.int32) if masks else None, (r % 2 == 0))

    pending = [rec(0, 4), rec(4, 4), rec(8, 3)]
    batch, rest = _take(pending, 6)
    assert batch[0].tolist() == [0, 1, 2, 3, 4, 5] and batch[4].shape == (6, 2) and batch[5].tolist() == [True, False] * 3
    assert [len(r[0]) for r in rest] == [2, 3] and rest[0][0].tolist() == [6, 7]
    batch, rest = _take(rest, 5)
    assert batch[0].tolist() == [6, 7, 8, 9, 10] and rest == []
    batch, rest = _take([rec(0, 3, masks=False)], 2)
    assert batch[4] is None and batch[3].tolist() == [1, 2] and len(rest[0][0]) == 1
