"""RLS with variable forgetting factor (BLUEROV2_AMPC::RLSFF, bluerov2_ampc.cpp:731-1004) and the AMPC variant of the
EKF model (bluerov2_ampc.cpp:658-696): the C restatement against an independent, literal numpy/list transcription."""
import numpy as np


class _RefAxis:
    """One axis of RLSFF() written the way the reference writes it (std::vector windows, Eigen expressions)."""

    def __init__(self):
        self.P = np.eye(4); self.theta = np.zeros(4); self.lam = 0.9          # :63-79
        self.err_n, self.err_d, self.F = [], [], 0.0

    def step(self, y, acc, vel):
        x = np.array([acc, vel, 1.0, vel * abs(vel)])
        e = y - x.dot(self.theta)
        self.err_n.append(e); self.err_d.append(e)
        if len(self.err_n) > 5: self.err_n.pop(0)
        if len(self.err_d) > 50: self.err_d.pop(0)
        var = []
        for w in (self.err_n, self.err_d):
            mean = sum(w, 0.0) / len(w)
            var.append(sum(((v - mean) ** 2 for v in w), 0.0) / len(w))
        with np.errstate(invalid="ignore", divide="ignore"):
            self.F = np.float64(var[0]) / np.float64(var[1])
        if self.F > 0.8:
            self.lam = self.lam - 0.01 if self.lam - 0.01 >= 0.5 else 0.5
        else:
            self.lam = self.lam + 0.01 if self.lam + 0.01 <= 1 else 1
        K = self.P @ x / (self.lam + x.dot(self.P @ x))
        self.theta = self.theta + K * e
        self.P = (self.P - np.outer(K, x) @ self.P) / self.lam


def _inputs(rng, nb):
    ex = rng.normal(0, 3, (nb, 18)); acc = rng.normal(0, 1, (nb, 6)); meas = rng.normal(0, 0.5, (nb, 12))
    return ex, acc, meas


def test_init(oracle):
    st = oracle.rls_init()
    assert st.shape == (4, 80)
    for a in range(4):
        assert np.array_equal(st[a, 0:4], np.zeros(4)) and np.array_equal(st[a, 4:20].reshape(4, 4), np.eye(4))
        assert st[a, 20] == 0.9 and st[a, 22] == 0 and st[a, 23] == 0


def test_rls_against_literal_transcription(oracle):
    rng = np.random.default_rng(5)
    nb, T = 3, 130                      # > 50 ticks: both windows wrap; lambda walks to its clamps
    st = oracle.rls_init(nb)
    ref = [[_RefAxis() for _ in range(4)] for _ in range(nb)]
    yi, ai, vi = (12, 13, 14, 17), (0, 1, 2, 5), (6, 7, 8, 11)
    lam_seen = set()
    for t in range(T):
        ex, acc, meas = _inputs(rng, nb)
        if t > 60:
            acc *= 0.01; meas *= 0.01; ex = 0.05 * ex + 2.0       # quiet phase: F drops, lambda climbs to 1
        p = oracle.rls_step_batch(st, ex, acc, meas, compensate=True)
        for b in range(nb):
            for a in range(4):
                r = ref[b][a]
                r.step(ex[b, yi[a]], acc[b, ai[a]], meas[b, vi[a]])
                s = st[b, a]
                assert np.allclose(s[0:4], r.theta, rtol=1e-9, atol=1e-10), (t, b, a)
                assert np.allclose(s[4:20].reshape(4, 4), r.P, rtol=1e-8, atol=1e-10), (t, b, a)
                assert abs(s[20] - r.lam) < 1e-12, (t, b, a, s[20], r.lam)
                assert (np.isnan(s[21]) and np.isnan(r.F)) or np.isclose(s[21], r.F, rtol=1e-9), (t, b, a)
                assert s[22] == len(r.err_n) and s[23] == len(r.err_d)
                assert np.allclose(s[24:24 + len(r.err_n)], r.err_n, rtol=1e-9, atol=1e-11)
                assert np.allclose(s[29:29 + len(r.err_d)], r.err_d, rtol=1e-9, atol=1e-11)
                lam_seen.add(round(float(s[20]), 2))
            # parameters handed to the OCP (:340-382)
            want = [ref[b][0].theta[2] / 0.032546960744430276, ref[b][1].theta[2] / 0.032546960744430276,
                    ref[b][2].theta[2] / 0.026546960744430276, ref[b][3].theta[2] / 0.026546960744430276]
            assert np.allclose(p[b, :4], want, rtol=1e-9, atol=1e-9)
            assert np.array_equal(p[b, 4:], [1.7182, 0, 5.468, 0.4006, -11.7391, -20, -31.8678, -5, -18.18, -21.66, -36.99, -1.55])
    assert 0.5 in lam_seen and 1.0 in lam_seen          # both clamps of the forgetting factor were exercised


def test_first_tick_nan_f_statistic_raises_lambda(oracle):
    st = oracle.rls_init(1)
    oracle.rls_step_batch(st, np.ones((1, 18)), np.ones((1, 6)), np.ones((1, 12)))
    assert np.isnan(st[0, :, 21]).all()                 # 0/0 with one sample in each window
    assert np.allclose(st[0, :, 20], 0.91)              # NaN > 0.8 is false -> lambda += 0.01


def test_no_compensation_leaves_hydrodynamic_parameters_untouched(oracle):
    st = oracle.rls_init(2)
    p = np.full((2, 16), 7.0)
    oracle.rls_step_batch(st, np.ones((2, 18)), np.ones((2, 6)), np.ones((2, 12)), compensate=False, p_out=p)
    assert np.array_equal(p[:, :4], np.zeros((2, 4))) and np.array_equal(p[:, 4:], np.full((2, 12), 7.0))


def test_ampc_filter_model_is_the_dob_model_without_damping(oracle):
    rng = np.random.default_rng(2)
    Dl = np.array([-11.7391, -20, -31.8678, -25, -44.9085, -5]); Dnl = np.array([-18.18, -21.66, -36.99, -1.55, -1.55, -1.55])
    m, Ix, Iy, Iz, ZG = 11.26, 0.3, 0.63, 0.58, 0.02
    am = [1.7182, 0, 5.468, 0, 1.2481, 0.4006]
    M = np.diag([m + am[0], m + am[1], m + am[2], Ix + am[3], Iy + am[4], Iz + am[5]])
    M[0, 4] = m * ZG; M[1, 3] = -m * ZG; M[3, 1] = -m * ZG; M[4, 0] = m * ZG
    iM = np.diag(np.linalg.inv(M))
    try:
        for _ in range(10):
            x = rng.uniform(-1, 1, 18); u = rng.uniform(-10, 10, 6)
            oracle.ekf_set_model(0)
            f0, h0 = oracle.ekf_f(x, u), oracle.ekf_h(x, u)
            oracle.ekf_set_model(1)
            f1, h1 = oracle.ekf_f(x, u), oracle.ekf_h(x, u)
            damp = Dl * x[6:12] + Dnl * np.abs(x[6:12]) * x[6:12]
            assert np.allclose(f0[6:12] - f1[6:12], iM * damp, rtol=1e-12, atol=1e-12)
            assert np.allclose(h1[12:] - h0[12:], damp, rtol=1e-12, atol=1e-12)
            assert np.array_equal(f0[:6], f1[:6]) and np.array_equal(h0[:12], h1[:12])
    finally:
        oracle.ekf_set_model(0)
