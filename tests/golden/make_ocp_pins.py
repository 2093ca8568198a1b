#!/usr/bin/env python
"""tests/golden/make_ocp_pins.py -- extract the problem data of the reference's fully resolved OCP description
(/root/reference/bluerov2_dobmpc/scripts/acados_ocp.json, written by generate_c_code.py next to the generated C) into
tests/golden/acados_ocp_pins.json.  It is the one reference-held fixture for the NLP/QP layer: weights, bounds, horizon,
integrator and QP-solver options.  Run where /root/reference exists; the fixture travels, the reference does not."""
import json
import os

import numpy as np

SRC = "/root/reference/bluerov2_dobmpc/scripts/acados_ocp.json"
HERE = os.path.dirname(os.path.abspath(__file__))


def extract(path=SRC):
    d = json.load(open(path))
    so, co, cn, dm = d["solver_options"], d["cost"], d["constraints"], d["dims"]
    W, W0, We = np.array(co["W"]), np.array(co["W_0"]), np.array(co["W_e"])
    Vx, Vu, Vxe = np.array(co["Vx"]), np.array(co["Vu"]), np.array(co["Vx_e"])

    def uniq(key):
        v = sorted(set(so[key]))
        assert len(v) == 1, (key, v)
        return v[0]

    return {
        "source": "bluerov2_dobmpc/scripts/acados_ocp.json",
        "dims": {k: dm[k] for k in ("N", "nx", "nu", "np", "ny", "ny_0", "ny_e", "nbu", "nbx", "nbx_0", "nbxe_0", "nbx_e", "ng", "nh", "ns")},
        "W_diag": np.diag(W).tolist(), "W_0_diag": np.diag(W0).tolist(), "W_e_diag": np.diag(We).tolist(),
        "W_offdiag_max": float(max(np.abs(M - np.diag(np.diag(M))).max() for M in (W, W0, We))),
        # y = [x; u]: Vx selects the state into rows 0..11, Vu the input into rows 12..15; terminal y = x
        "Vx_is_selector": bool(np.array_equal(Vx, np.vstack([np.eye(12), np.zeros((4, 12))]))),
        "Vu_is_selector": bool(np.array_equal(Vu, np.vstack([np.zeros((12, 4)), np.eye(4)]))),
        "Vx_e_is_identity": bool(np.array_equal(Vxe, np.eye(12))),
        "cost_type": [co["cost_type_0"], co["cost_type"], co["cost_type_e"]],
        "yref": co["yref"], "yref_e": co["yref_e"],
        "lbu": cn["lbu"], "ubu": cn["ubu"], "idxbu": cn["idxbu"],
        "lbx_0": cn["lbx_0"], "ubx_0": cn["ubx_0"], "idxbx_0": cn["idxbx_0"], "idxbxe_0": cn["idxbxe_0"],
        "constr_type": cn["constr_type"],
        "parameter_values": d["parameter_values"],
        "tf": so["tf"], "time_step": uniq("time_steps"), "n_time_steps": len(so["time_steps"]), "Tsim": so["Tsim"],
        "integrator_type": so["integrator_type"], "sim_method_num_stages": uniq("sim_method_num_stages"),
        "sim_method_num_steps": uniq("sim_method_num_steps"),
        "nlp_solver_type": so["nlp_solver_type"], "hessian_approx": so["hessian_approx"], "qp_solver": so["qp_solver"],
        "hpipm_mode": so["hpipm_mode"], "qp_solver_iter_max": so["qp_solver_iter_max"],
        "qp_solver_warm_start": so["qp_solver_warm_start"], "globalization": so["globalization"],
        "nlp_solver_step_length": so["nlp_solver_step_length"], "levenberg_marquardt": so["levenberg_marquardt"],
        "full_step_dual": so["full_step_dual"], "print_level": so["print_level"],
        "model_name": d["model"]["name"],
    }


if __name__ == "__main__":
    out = os.path.join(HERE, "acados_ocp_pins.json")
    json.dump(extract(), open(out, "w"), indent=1, sort_keys=True)
    print(out)
