import os

import numpy as np
import pytest

from bluerov2_b200 import traj

REF = "/root/reference/bluerov2_path/config/traj"


def test_shapes_and_quirks():
    c = traj.circle()
    assert c.shape == (4801, 16)
    assert np.all(c[:, 6] == 1.5) and np.all(c[:, 7] == 1.498945)     # circle.py:45-46 flatten-index slip
    assert np.all(c[:, 14] == 57.5)                                    # circle.py:55, outside the +-50 box
    l = traj.lemniscate()
    assert l.shape == (1201, 16) and np.all(l[:, 2] == -20.0)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_bit_identical_to_reference_files():
    assert np.array_equal(traj.circle(), np.loadtxt(os.path.join(REF, "circle.txt")))
    assert np.array_equal(traj.lemniscate(), np.loadtxt(os.path.join(REF, "lemniscate.txt")))


def test_checksums_of_the_reference_files():
    """sha256 of the parsed reference files (bluerov2_path/config/traj/{circle,lemniscate}.txt), recorded with the reference
    present: pins the generators on machines that do not have /root/reference (the GPU box)."""
    import hashlib
    want = {"circle": "add3016bd03d34de26a424635db2aca70c5932b6ef0dcbbd52f57bcaed153212",
            "lemniscate": "4b10892f5bcb53486e35b8b43fbd019daa405bc7a6287ed0e5d09a13c667eba8"}
    for name, fn in (("circle", traj.circle), ("lemniscate", traj.lemniscate)):
        assert hashlib.sha256(np.ascontiguousarray(fn() + 0.0).tobytes())   # + 0.0: one -0.0 in the file.hexdigest() == want[name], name


def test_window_clamps_to_last_row():
    """ref_cb (bluerov2_dob.cpp:218-265): all three branches."""
    t = np.arange(10 * 16, dtype=float).reshape(10, 16)
    w = traj.window(t, 2, 4)
    assert np.array_equal(w, t[2:7])
    w = traj.window(t, 7, 4)          # partially past the end
    assert np.array_equal(w[:3], t[7:10]) and np.all(w[3:] == t[9])
    w = traj.window(t, 40, 4)         # entirely past the end
    assert np.all(w == t[9])
    wb = traj.window_batch(t, np.array([2, 7, 40]), 4)
    assert np.array_equal(wb[0], traj.window(t, 2, 4)) and np.array_equal(wb[1], traj.window(t, 7, 4))


def test_ctrl_node_glue():
    """parameter / reference fill of the consolidated node (src/ctrller/mpc.cpp:139-187, 199-262)"""
    from bluerov2_b200 import workloads as wl
    d = np.array([[4.0, -2.0, 1.0], [0.0, 0.5, -3.0]])
    p = wl.ctrl_params(d, dompc=True)
    assert np.allclose(p[:, 0], d[:, 0] / 0.032546960744430276) and np.allclose(p[:, 1], d[:, 1] / 0.032546960744430276)
    assert np.allclose(p[:, 2], d[:, 2] / 0.026546960744430276) and np.all(p[:, 3] == 0.0)
    assert np.array_equal(p[:, 4:], np.tile(wl.NOMINAL_P[4:], (2, 1)))
    assert np.array_equal(wl.ctrl_params(d, dompc=False), np.tile(wl.NOMINAL_P, (2, 1)))
    y = wl.ctrl_yref(np.ones((3, 5, 12)))
    assert y.shape == (3, 5, 16) and np.all(y[..., :12] == 1.0) and np.all(y[..., 12:] == 0.0)
