"""CPU: FortAttackBatch refuses host streams of the wrong shape / type before any pointer reaches the library
(fa_step_many_host writes T steps through raw host pointers)."""
from importlib import import_module

import pytest
import torch

be = import_module("emergent-multiagent-strategies_b200.batched_env")


def _fake(A=6, E=10, dtype=torch.float32):
    b = be.FortAttackBatch.__new__(be.FortAttackBatch)       # no device needed for the argument check
    b.A, b.E, b.dtype = A, E, dtype
    return b


def _streams(T=4, A=6, E=10, dtype=torch.float32):
    return [torch.zeros(T, A, E, dtype=torch.int32), torch.zeros(T, A, E, 6, dtype=dtype), torch.zeros(T, A, E, dtype=dtype),
            torch.zeros(T, E, dtype=torch.uint8), torch.zeros(T, E, dtype=torch.uint8)]


def test_accepts_the_documented_layout_and_optional_outputs():
    b = _fake()
    assert b._check_host_streams(*_streams()) == 4
    s = _streams()
    s[1] = None                                              # outputs may be skipped
    s[4] = None
    assert b._check_host_streams(*s) == 4
    assert _fake(dtype=torch.float64)._check_host_streams(*_streams(dtype=torch.float64)) == 4


@pytest.mark.parametrize("which,bad", [
    (0, torch.zeros(4, 6, 9, dtype=torch.int32)), (0, torch.zeros(4, 6, 10, dtype=torch.int64)), (0, torch.zeros(6, 10, dtype=torch.int32)),
    (0, torch.zeros(0, 6, 10, dtype=torch.int32)), (1, torch.zeros(3, 6, 10, 6)), (1, torch.zeros(4, 6, 10, 6, dtype=torch.float64)),
    (1, torch.zeros(4, 6, 10, 12)[..., ::2]), (2, torch.zeros(4, 6, 10, 1)), (3, torch.zeros(4, 10)), (4, torch.zeros(5, 10, dtype=torch.uint8))])
def test_rejects_wrong_streams(which, bad):
    s = _streams()
    s[which] = bad
    with pytest.raises(ValueError):
        _fake()._check_host_streams(*s)
    s = _streams()
    s[0] = None
    with pytest.raises((ValueError, AttributeError)):
        _fake()._check_host_streams(*s)


def test_step_host_rejects_wrong_buffers_before_touching_the_library():
    b = _fake()
    ok = [torch.zeros(6, 10, dtype=torch.int32), torch.zeros(6, 10, 6), torch.zeros(6, 10), torch.zeros(10, dtype=torch.uint8),
          torch.zeros(10, dtype=torch.uint8)]
    for which, bad in ((0, torch.zeros(6, 10)), (1, torch.zeros(6, 10, 5)), (2, torch.zeros(10, 6).t()), (3, torch.zeros(10)),
                       (4, torch.zeros(11, dtype=torch.uint8))):
        args = list(ok)
        args[which] = bad
        with pytest.raises(ValueError):
            b.step_host(*args)          # raises in the argument check: the fake object has no library handle at all
