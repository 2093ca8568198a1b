"""Host-argument marshalling cache of bluerov2_b200.solver (no GPU needed)."""
import numpy as np
import pytest

from bluerov2_b200.solver import _HostArgs


def test_cached_pointer_and_conversions():
    h = _HostArgs(cap=4)
    a = np.zeros((4, 12))
    keep, p = h.get(a, (4, 12))
    assert keep is a and p.value == a.ctypes.data
    assert h.get(a, (4, 12))[1] is p                       # second call: remembered
    b = np.ones((4, 12), dtype=np.float32)                 # wrong dtype: converted for this call only, never cached
    keep, p = h.get(b, (4, 12))
    assert keep is not b and keep.dtype == np.float64 and p.value == keep.ctypes.data and np.all(keep == 1.0)
    c = np.arange(96.0).reshape(8, 12)[::2]                # non-contiguous view
    keep, p = h.get(c, (4, 12))
    assert keep.flags.c_contiguous and np.array_equal(keep, c)
    keep, p = h.get([[0.0] * 12] * 4, (4, 12))             # nested list
    assert keep.shape == (4, 12)
    with pytest.raises(ValueError):
        h.get(np.zeros((3, 12)), (4, 12))
    lines = np.zeros(4, dtype=np.int32)
    assert h.get(lines, (4,), np.int32)[0] is lines


def test_outputs_are_never_converted():
    h = _HostArgs()
    with pytest.raises(ValueError):
        h.get(np.zeros((4, 6), dtype=np.float32), (4, 6), out=True)
    with pytest.raises(ValueError):
        h.get(np.zeros((8, 6))[::2], (4, 6), out=True)
    ro = np.zeros((4, 6)); ro.flags.writeable = False
    with pytest.raises(ValueError):
        h.get(ro, (4, 6), out=True)
    ok = np.zeros((4, 6))
    assert h.get(ok, (4, 6), out=True)[0] is ok


def test_capacity_bound_and_identity_check():
    h = _HostArgs(cap=2)
    arrs = [np.zeros((2, 4)) for _ in range(5)]
    for a in arrs:
        assert h.get(a, (2, 4))[1].value == a.ctypes.data
    assert len(h._d) <= 2
    # same shape asked under another layout contract -> re-validated, not served from the cache
    a = np.zeros((2, 4))
    h.get(a, (2, 4))
    with pytest.raises(ValueError):
        h.get(a, (4, 2))
