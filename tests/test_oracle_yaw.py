"""Continuous-yaw accumulator of BLUEROV2_DOB::solve (bluerov2_dob.cpp:272-304): the C restatement against a literal numpy
transcription that keeps the reference's types (pre_yaw / yaw_sum / yaw_diff float32, psi float64)."""
import numpy as np


def _wrap(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


def _ref_step(pre_yaw, yaw_sum, psi):
    """types as declared in bluerov2_dob.h:79-83,234-236"""
    pre_yaw, yaw_sum, psi = np.float32(pre_yaw), np.float32(yaw_sum), np.float64(psi)
    two_pi = np.float64(2 * np.pi)
    if pre_yaw >= 0 and psi >= 0:
        d = np.float32(psi - np.float64(pre_yaw))
    elif pre_yaw >= 0 and psi < 0:
        if two_pi + psi - np.float64(pre_yaw) >= np.float64(pre_yaw) + abs(psi):
            d = np.float32(-(np.float64(pre_yaw) + abs(psi)))
        else:
            d = np.float32(two_pi + psi - np.float64(pre_yaw))
    elif pre_yaw < 0 and psi >= 0:
        if two_pi - psi + np.float64(pre_yaw) >= np.float64(abs(pre_yaw)) + psi:
            d = np.float32(np.float64(abs(pre_yaw)) + psi)
        else:
            d = np.float32(-(two_pi - psi + np.float64(pre_yaw)))
    else:
        d = np.float32(psi - np.float64(pre_yaw))
    yaw_sum = np.float32(yaw_sum + d)
    return np.float32(psi), yaw_sum


def test_against_literal_transcription_and_true_angle(oracle):
    rng = np.random.default_rng(3)
    nb, T = 16, 400
    true = rng.uniform(-0.5, 0.5, nb)                   # continuous yaw, starts near 0 like the node's accumulators
    rate = rng.uniform(-0.6, 0.6, nb)                   # rad per tick: several full turns in both directions
    st = np.zeros((nb, 2), dtype=np.float32)
    ref = [(np.float32(0), np.float32(0)) for _ in range(nb)]
    peak = 0.0
    for t in range(T):
        true = true + rate + rng.normal(0, 0.05, nb)
        if t % 97 == 0:
            rate = -rate
        peak = max(peak, float(np.abs(true).max()))
        psi = _wrap(true)
        out = oracle.yaw_unwrap_batch(st, psi)
        for b in range(nb):
            ref[b] = _ref_step(ref[b][0], ref[b][1], psi[b])
            assert st[b, 0] == ref[b][0] and st[b, 1] == ref[b][1], (t, b)       # bit for bit
            assert out[b] == np.float64(ref[b][1])
        # and it IS the continuous angle, up to the float32 accumulation the reference commits to
        assert np.abs(out - true).max() < 1e-3 * max(1.0, np.abs(true).max()), (t, np.abs(out - true).max())
    assert peak > 4 * np.pi                             # the walk did wrap many times


def test_branch_table(oracle):
    """the four sign cases and both sides of the half-turn decision"""
    cases = [(0.5, 1.0), (3.0, -3.0), (1.0, -1.0), (-3.0, 3.0), (-1.0, 1.0), (-0.5, -1.0), (0.0, -0.0), (3.1, -3.1)]
    for pre, psi in cases:
        st = np.array([[pre, 10.0]], dtype=np.float32)
        out = oracle.yaw_unwrap_batch(st, np.array([psi]))
        want = _ref_step(np.float32(pre), np.float32(10.0), psi)
        assert st[0, 0] == want[0] and st[0, 1] == want[1] and out[0] == np.float64(want[1]), (pre, psi)
        d = float(st[0, 1]) - 10.0
        assert abs(d - _wrap(psi - pre)) < 1e-5, (pre, psi, d)      # shortest signed difference
