"""Pin of the baked problem data on the one reference-held fixture of the NLP/QP layer: scripts/acados_ocp.json, the fully
resolved OCP description generate_c_code.py writes next to the generated C.  tests/golden/acados_ocp_pins.json is its extract
(tests/golden/make_ocp_pins.py); where /root/reference exists the extract is re-derived and compared first.

Checked against it: the defaults table of the product library (br2_get_ocp_defaults: what br2_batch_create and the acados-ABI
shim bake), the macros of include/acados_solver_bluerov2.h, and the constants of the oracle (oracle/oracle.py)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from tests.conftest import GOLDEN, REFERENCE, ROOT


@pytest.fixture(scope="module")
def pins():
    return json.load(open(os.path.join(GOLDEN, "acados_ocp_pins.json")))


class Defaults(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("N", "nx", "nu", "np", "ny", "ny_e", "nbu", "nbx0", "nbxe0", "qp_iter_max", "qp_warm_start",
                                       "erk_stages", "erk_steps")] + \
               [("Tf", C.c_double), ("W", C.c_double * 16), ("We", C.c_double * 12), ("lbu", C.c_double * 4), ("ubu", C.c_double * 4),
                ("x_init", C.c_double * 12)]


@pytest.fixture(scope="module")
def defaults():
    from bluerov2_b200 import solver
    L = solver.load_library()
    L.br2_get_ocp_defaults.argtypes = [C.POINTER(Defaults)]
    L.br2_get_ocp_defaults.restype = None
    d = Defaults()
    L.br2_get_ocp_defaults(C.byref(d))
    return d


def test_fixture_is_the_reference_extract(pins):
    src = os.path.join(REFERENCE, "bluerov2_dobmpc", "scripts", "acados_ocp.json")
    if not os.path.exists(src):
        pytest.skip("/root/reference not present: the committed extract stands")
    from tests.golden.make_ocp_pins import extract
    assert extract(src) == pins


def test_problem_structure_is_what_the_engine_assumes(pins):
    """the engine hard-codes: diagonal weights, y = [x; u] (terminal y = x), NONLINEAR_LS everywhere, BGH box on all 4 inputs,
    all 12 stage-0 state bounds equalities, no other constraints, SQP_RTI + Gauss-Newton + ERK4 x 1 step, fixed full steps,
    no Levenberg-Marquardt term, cold-started QP"""
    assert pins["W_offdiag_max"] == 0.0
    assert pins["Vx_is_selector"] and pins["Vu_is_selector"] and pins["Vx_e_is_identity"]
    assert pins["cost_type"] == ["NONLINEAR_LS"] * 3
    assert pins["W_0_diag"] == pins["W_diag"]
    assert pins["constr_type"] == "BGH" and pins["idxbu"] == [0, 1, 2, 3]
    assert pins["idxbx_0"] == list(range(12)) and pins["idxbxe_0"] == list(range(12))
    d = pins["dims"]
    assert (d["nbx"], d["nbx_e"], d["ng"], d["nh"], d["ns"]) == (0, 0, 0, 0, 0)
    assert pins["nlp_solver_type"] == "SQP_RTI" and pins["hessian_approx"] == "GAUSS_NEWTON"
    assert pins["integrator_type"] == "ERK" and pins["globalization"] == "FIXED_STEP"
    assert pins["nlp_solver_step_length"] == 1.0 and pins["levenberg_marquardt"] == 0.0 and pins["full_step_dual"] == 0
    assert pins["qp_solver"] == "FULL_CONDENSING_HPIPM" and pins["hpipm_mode"] == "BALANCE"
    assert pins["model_name"] == "bluerov2"
    assert pins["yref"] == [0.0] * 16 and pins["yref_e"] == [0.0] * 12 and pins["parameter_values"] == [0.0] * 16


def test_library_defaults_match_acados_ocp_json(pins, defaults):
    d, dm = defaults, pins["dims"]
    assert (d.N, d.nx, d.nu, d.np, d.ny, d.ny_e) == (dm["N"], dm["nx"], dm["nu"], dm["np"], dm["ny"], dm["ny_e"])
    assert (d.nbu, d.nbx0, d.nbxe0) == (dm["nbu"], dm["nbx_0"], dm["nbxe_0"])
    assert list(d.W) == pins["W_diag"] and list(d.We) == [float(v) for v in pins["W_e_diag"]]
    assert list(d.lbu) == [float(v) for v in pins["lbu"]] and list(d.ubu) == [float(v) for v in pins["ubu"]]
    assert list(d.x_init) == pins["lbx_0"] == pins["ubx_0"]
    assert d.Tf == pins["tf"] and pins["n_time_steps"] == d.N
    assert d.Tf / d.N == pins["time_step"] == pins["Tsim"]
    assert d.qp_iter_max == pins["qp_solver_iter_max"] and d.qp_warm_start == pins["qp_solver_warm_start"]
    assert d.erk_stages == pins["sim_method_num_stages"] and d.erk_steps == pins["sim_method_num_steps"]


def test_header_macros_match_acados_ocp_json(pins):
    txt = open(os.path.join(ROOT, "include", "acados_solver_bluerov2.h")).read()
    mac = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+BLUEROV2_(\w+)\s+(\d+)", txt)}
    dm = pins["dims"]
    want = {"NX": dm["nx"], "NU": dm["nu"], "NP": dm["np"], "N": dm["N"], "NY": dm["ny"], "NY0": dm["ny_0"], "NYN": dm["ny_e"],
            "NBU": dm["nbu"], "NBX": dm["nbx"], "NBX0": dm["nbx_0"], "NBXN": dm["nbx_e"], "NG": dm["ng"], "NH": dm["nh"], "NS": dm["ns"]}
    for k, v in want.items():
        assert mac[k] == v, (k, mac.get(k), v)


def test_oracle_constants_match_acados_ocp_json(pins):
    from oracle import W_DEFAULT, WE_DEFAULT, LBU, UBU, X_INIT
    assert np.array_equal(W_DEFAULT, pins["W_diag"]) and np.array_equal(WE_DEFAULT, pins["W_e_diag"])
    assert np.array_equal(LBU, pins["lbu"]) and np.array_equal(UBU, pins["ubu"]) and np.array_equal(X_INIT, pins["lbx_0"])
