"""bluerov2_b200/csrc/fast_trig.h: the branch-free sincos the device code uses, built for the host with g++ and checked against
long-double libm (the device build differs only in how the integer is read out of the rounding constant)."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("fast_trig")
    src = d / "ft.cpp"
    src.write_text('#include "fast_trig.h"\nextern "C" void ft_sincos(const double* x, double* s, double* c, int n)'
                   ' { for (int i = 0; i < n; i++) br2_sincos(x[i], s + i, c + i); }\n')
    so = d / "ft.so"
    subprocess.check_call(["g++", "-O2", "-mfma", "-shared", "-fPIC", "-I", os.path.join(ROOT, "bluerov2_b200", "csrc"), str(src), "-o", str(so)])
    return C.CDLL(str(so))


def run(lib, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s, c = np.empty_like(x), np.empty_like(x)
    lib.ft_sincos(x.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), C.c_int(x.size))
    return s, c


@pytest.mark.parametrize("span", [np.pi, 100.0, 1e5])
def test_sincos_within_two_ulp(lib, span):
    x = np.random.default_rng(0).uniform(-span, span, 400_000)
    s, c = run(lib, x)
    xl = x.astype(np.longdouble)
    rs, rc = np.sin(xl), np.cos(xl)
    assert (np.abs(s - rs) / np.spacing(np.abs(rs.astype(np.float64)))).max() < 2.0
    assert (np.abs(c - rc) / np.spacing(np.abs(rc.astype(np.float64)))).max() < 2.0


def test_sincos_near_multiples_of_half_pi_and_specials(lib):
    rng = np.random.default_rng(1)
    x = (np.arange(-2000, 2000)[:, None] * (np.pi / 2) + rng.uniform(-1e-6, 1e-6, (4000, 20))).ravel()
    s, c = run(lib, x)
    xl = x.astype(np.longdouble)
    assert np.abs(s - np.sin(xl)).max() < 1.2e-16 and np.abs(c - np.cos(xl)).max() < 1.2e-16
    s, c = run(lib, np.array([np.nan, np.inf, -np.inf, 0.0]))
    assert np.isnan(s[:3]).all() and np.isnan(c[:3]).all() and s[3] == 0.0 and c[3] == 1.0
