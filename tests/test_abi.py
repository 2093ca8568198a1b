"""The drop-in boundary: include/*.h vs the built library, and the reference's own harness linked against it."""
import ctypes as C
import os
import re
import struct
import subprocess
import sys

import numpy as np
import pytest

from tests.conftest import ROOT, REFERENCE

INCLUDE = os.path.join(ROOT, "include")
ABI_BIN = os.path.join(ROOT, "tests", "abi", "_bin")


@pytest.fixture(scope="module")
def lib_path():
    from bluerov2_b200 import build
    return build.build_library()


def _declared_functions():
    """every function prototype in include/**/*.h after preprocessing (some headers stamp their prototypes out of a macro);
    only text that originates from files under include/ is scanned (line markers), so libc prototypes do not count"""
    names = set()
    proto = re.compile(r"^\s*(?:[A-Za-z_]\w*[\s\*]+)+([A-Za-z_]\w*)\s*\(", re.M)     # type words / stars, then name(
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    for r, _, fs in os.walk(INCLUDE):
        for f in fs:
            if not f.endswith(".h"):
                continue
            out = subprocess.run([cc, "-E", "-I", INCLUDE, "-I", os.path.join(INCLUDE, "blasfeo", "include"), os.path.join(r, f)],
                                 capture_output=True, text=True)
            assert out.returncode == 0, (f, out.stderr[-2000:])
            keep, ours = [], False
            for line in out.stdout.splitlines():
                if line.startswith("# "):
                    m = re.match(r'# \d+ "([^"]*)"', line)
                    ours = bool(m) and os.path.abspath(m.group(1)).startswith(INCLUDE)
                    continue
                if ours:
                    keep.append(line)
            src = "\n".join(keep)
            src = re.sub(r"typedef\s+struct[^{;]*\{.*?\}\s*[A-Za-z_0-9]*\s*;", "", src, flags=re.S)   # drop struct bodies
            src = re.sub(r"__attribute__\s*\(\(.*?\)\)\)?", "", src)                                    # export attributes
            src = src.replace(";", ";\n")                                                            # one declaration per line
            for m in proto.finditer(src):
                n = m.group(1)
                if n not in ("defined", "sizeof", "if", "visibility", "__attribute__", "__declspec"):
                    names.add(n)
    return names


def test_every_declared_symbol_is_exported(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    declared = _declared_functions()
    assert len(declared) >= 60, sorted(declared)
    must = {"bluerov2_acados_create_capsule", "bluerov2_acados_free_capsule", "bluerov2_acados_create", "bluerov2_acados_reset",
            "bluerov2_acados_create_with_discretization", "bluerov2_acados_update_time_steps",
            "bluerov2_acados_update_qp_solver_cond_N", "bluerov2_acados_update_params", "bluerov2_acados_update_params_sparse",
            "bluerov2_acados_solve", "bluerov2_acados_free", "bluerov2_acados_print_stats", "bluerov2_acados_custom_update",
            "bluerov2_acados_get_nlp_in", "bluerov2_acados_get_nlp_out", "bluerov2_acados_get_sens_out",
            "bluerov2_acados_get_nlp_solver", "bluerov2_acados_get_nlp_config", "bluerov2_acados_get_nlp_opts",
            "bluerov2_acados_get_nlp_dims", "bluerov2_acados_get_nlp_plan",
            "ocp_nlp_constraints_model_set", "ocp_nlp_cost_model_set", "ocp_nlp_out_get", "ocp_nlp_out_set", "ocp_nlp_get",
            "ocp_nlp_solver_opts_set", "d_print_exp_tran_mat", "br2_batch_create", "br2_batch_solve_device",
            "br2_batch_solve_host", "br2_batch_ekf_device"}
    assert must <= declared, must - declared
    missing = declared - exported
    assert not missing, f"declared in include/ but not exported by {os.path.basename(lib_path)}: {sorted(missing)}"


def test_library_loads_and_reports_version(lib_path):
    from bluerov2_b200 import solver
    L = solver.load_library()
    assert b"sm_100a" in L.br2_version()
    assert L.br2_device_count() >= 0


def test_link_shims_exist(lib_path):
    d = os.path.dirname(lib_path)
    for s in ("libacados.so", "libhpipm.so", "libblasfeo.so"):
        out = subprocess.run(["readelf", "-d", os.path.join(d, s)], capture_output=True, text=True, check=True).stdout
        assert "libacados_ocp_solver_bluerov2.so" in out


def test_capsule_layout_and_macros(lib_path, tmp_path):
    """macros and capsule members the callers use (bluerov2_dob.h:68-77,166; bluerov2_dob.cpp:320,371,384-388)"""
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "acados_solver_bluerov2.h"
int main(void) {
    printf("%d %d %d %d %d %d %d\n", BLUEROV2_NX, BLUEROV2_NU, BLUEROV2_NP, BLUEROV2_NY, BLUEROV2_NYN, BLUEROV2_N, BLUEROV2_NBX0);
    printf("%d\n", (int)ACADOS_SUCCESS);
    printf("%zu %zu %zu\n", offsetof(bluerov2_solver_capsule, nlp_in), offsetof(bluerov2_solver_capsule, nlp_out),
           offsetof(bluerov2_solver_capsule, nlp_solver));
    ocp_nlp_out o; o.inf_norm_res = 1.5; ocp_nlp_dims d; d.N = 3;
    printf("%g %d\n", o.inf_norm_res, d.N);
    bluerov2_solver_capsule *c = bluerov2_acados_create_capsule();   /* must not need CUDA (bluerov2_dob.h:168) */
    printf("%d\n", c != NULL);
    return bluerov2_acados_free_capsule(c);
}''')
    exe = tmp_path / "t"
    d = os.path.dirname(lib_path)
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", INCLUDE, str(src), "-o", str(exe), "-L", d, "-lacados_ocp_solver_bluerov2",
                    "-Wl,-rpath," + d], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0] == "12 4 16 16 12 80 12"
    assert out[1] == "0"
    assert out[2] == "0 8 24"
    assert out[3] == "1.5 3"
    assert out[4] == "1"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="needs /root/reference")
def test_reference_main_compiles_and_links_unmodified(lib_path):
    """c_generated_code/main_bluerov2.c, compiled where it lies, against our headers and libraries only"""
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "abi"), "all"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert os.path.exists(os.path.join(ABI_BIN, "main_bluerov2"))
    nm = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(ABI_BIN, "main_bluerov2")], capture_output=True, text=True).stdout
    for sym in ("bluerov2_acados_create_with_discretization", "ocp_nlp_out_get", "d_print_exp_tran_mat", "bluerov2_acados_solve"):
        assert sym in nm


def test_model_functions_match_casadi_golden(lib_path, golden):
    """bluerov2_expl_ode_fun / _vde_forw / _vde_adj exported by the product library (csrc/model.cuh evaluated on the
    host: the same source the CUDA kernels compile for the device) against the reference's CasADi outputs."""
    L = C.CDLL(lib_path)
    g = golden["casadi_vde"]
    DP = C.POINTER(C.c_double)
    sig = [C.POINTER(DP), C.POINTER(DP), C.c_void_p, C.c_void_p, C.c_void_p]

    def call(name, ins, shapes):
        fn = getattr(L, name); fn.argtypes = sig; fn.restype = C.c_int
        ins = [np.ascontiguousarray(a, dtype=np.float64) for a in ins]
        outs = [np.zeros(s) for s in shapes]
        arg = (DP * len(ins))(*[a.ctypes.data_as(DP) for a in ins])
        res = (DP * len(outs))(*[a.ctypes.data_as(DP) for a in outs])
        assert fn(arg, res, None, None, None) == 0
        return outs

    for i in range(g["x"].shape[0]):
        f, dSx, dSu = call("bluerov2_expl_vde_forw", [g["x"][i], g["Sx_cm"][i], g["Su_cm"][i], g["u"][i], g["p"][i]], [(12,), (144,), (48,)])
        sc = max(1.0, np.abs(g["f"][i]).max())
        assert np.abs(f - g["f"][i]).max() < 1e-12 * sc
        assert np.abs(dSx - g["dSx_cm"][i]).max() < 1e-11 * max(1.0, np.abs(g["dSx_cm"][i]).max())
        assert np.abs(dSu - g["dSu_cm"][i]).max() < 1e-11 * max(1.0, np.abs(g["dSu_cm"][i]).max())
        (fo,) = call("bluerov2_expl_ode_fun", [g["x"][i], g["u"][i], g["p"][i]], [(12,)])
        assert np.abs(fo - g["f_ode"][i]).max() < 1e-12 * sc
        (adj,) = call("bluerov2_expl_vde_adj", [g["x"][i], g["lam"][i], g["u"][i], g["p"][i]], [(16,)])
        assert np.abs(adj - g["adj"][i]).max() < 1e-11 * max(1.0, np.abs(g["adj"][i]).max())


def test_create_without_gpu_fails_loudly(lib_path):
    """no CPU fallback: on a box without a CUDA device, create prints the reason and exits 1 (like a failed
    ocp_nlp_precompute, acados_solver_bluerov2.c:725-728)"""
    from bluerov2_b200 import solver
    if solver.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    exe = os.path.join(ABI_BIN, "dob_tick")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "abi"), os.path.join(ABI_BIN, "dob_tick")], check=True)
    inp = os.path.join(ROOT, "tests", "abi", "_bin", "empty.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("ii", 10, 0))
    out = subprocess.run([exe, inp, inp + ".out"], capture_output=True, text=True)
    assert out.returncode == 1
    assert "no CUDA device" in out.stderr and "no CPU path" in out.stderr
    with pytest.raises(solver.SolverError):
        solver.BatchSolver(4, 10)


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_reference_main_runs(lib_path):
    exe = os.path.join(ABI_BIN, "main_bluerov2")
    if not os.path.exists(exe):
        pytest.skip("tests/abi/_bin/main_bluerov2 not prebuilt (needs /root/reference at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "bluerov2_acados_solve(): SUCCESS!" in out.stdout
    assert "--- utraj ---" in out.stdout and "SQP iterations  1" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("N", [80, 20])
def test_node_call_sequence_matches_oracle(lib_path, oracle, N, tmp_path):
    """BLUEROV2_DOB::solve's call sequence from C (tests/abi/dob_tick.c), 6 closed-loop ticks, against the oracle.
    N = 80 is the generated default (bluerov2_acados_create); N = 20 goes through create_with_discretization."""
    from bluerov2_b200 import traj, workloads as wl
    exe = os.path.join(ABI_BIN, "dob_tick")
    assert os.path.exists(exe), "run `make -C tests/abi` (done by __graft_entry__.build())"
    T = 6
    w = wl.tracking_batch(1, N, seed=3, pos_spread=2.5)
    p = w["p"][0].copy(); p[:4] = [4.0, -3.0, 2.0, 0.3]
    Ts = wl.time_steps(N)
    X, U = w["X"][0].copy(), w["U"][0].copy()
    x0, line = w["x0"][0].copy(), int(w["lines"][0])
    ticks, want = [], []
    for t in range(T):
        yref = traj.window(w["traj"], line + t, N)
        ticks.append((x0.copy(), yref))
        st, _ = oracle.rti_step(Ts, x0, yref, p, X, U)
        assert st == 0
        want.append(U[0].copy())
        x0 = oracle.erk4(x0, U[0], p, 0.05)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("ii", N, T))
        for x0t, yref in ticks:
            f.write(x0t.tobytes()); f.write(p.tobytes()); f.write(np.ascontiguousarray(yref).tobytes())
    out = subprocess.run([exe, str(fin), str(fout)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = open(fout, "rb").read()
    rec = 8 + 8 * 6
    for t in range(T):
        st = struct.unpack_from("i", raw, t * rec)[0]
        vals = np.frombuffer(raw, dtype=np.float64, count=6, offset=t * rec + 8)
        assert st == 0
        assert np.abs(vals[:4] - want[t]).max() < 1e-5, (t, vals[:4], want[t])
        assert 0 < vals[4] < 5.0 and np.isfinite(vals[5])
    tail = np.frombuffer(raw, dtype=np.float64, offset=T * rec)
    Xg, Ug = tail[:(N + 1) * 12].reshape(N + 1, 12), tail[(N + 1) * 12:].reshape(N, 4)
    assert np.abs(Ug - U).max() < 1e-5 and np.abs(Xg - X).max() < 1e-5


@pytest.mark.parametrize("lang", ["c", "c++"])
def test_public_headers_compile_standalone(tmp_path, lang):
    """every header a consumer includes (the list of bluerov2_dob.h:25-35 plus the batched API) compiles on its own as C99
    and as C++, warnings as errors -- no CUDA, no torch types in the signatures"""
    hdrs = ["acados/utils/print.h", "acados/utils/types.h", "acados/utils/math.h", "acados_c/ocp_nlp_interface.h",
            "acados_c/external_function_interface.h", "acados/ocp_nlp/ocp_nlp_constraints_bgh.h", "acados/ocp_nlp/ocp_nlp_cost_ls.h",
            "blasfeo/include/blasfeo_d_aux.h", "blasfeo/include/blasfeo_d_aux_ext_dep.h", "bluerov2_model/bluerov2_model.h",
            "bluerov2_cost/bluerov2_cost.h", "bluerov2_constraints/bluerov2_constraints.h", "acados_solver_bluerov2.h",
            "bluerov2_b200.h"]
    ext = "c" if lang == "c" else "cpp"
    src = tmp_path / f"hdr.{ext}"
    src.write_text("".join(f'#include "{h}"\n' for h in hdrs) + "int main(void) { return (int)sizeof(bluerov2_solver_capsule) == 0; }\n")
    cc = ("/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc") if lang == "c" else ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")
    std = "-std=c99" if lang == "c" else "-std=c++14"
    out = subprocess.run([cc, std, "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", INCLUDE, str(src)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]


# ---------------------------------------------------------------------------------------------------------
# behaviour of the generated-ABI functions around the solve, tests/abi/a10_behaviour.c
def _a10(mode, timeout=120):
    exe = os.path.join(ABI_BIN, "a10_behaviour")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "abi"), exe], check=True)
    return subprocess.run([exe, mode], capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("mode,needle", [
    ("cond_N", "acados_update_qp_solver_cond_N() failed, since no partial condensing solver is used!"),   # gen.c:788-794
    ("params_np", "trying to set 15 parameters for external functions. External function has 16 parameters. Exiting."),   # :839-844
    ("sparse_np", "trying to set 17 parameters for external functions. External function has 16 parameters. Exiting."),   # :890-896
])
def test_misuse_prints_and_exits_1(lib_path, mode, needle):
    """fatal misuse -> the reference's message and exit(1), before anything needs a CUDA device"""
    out = _a10(mode)
    assert out.returncode == 1, (out.returncode, out.stdout, out.stderr)
    assert needle in out.stdout and "returned" not in out.stdout


def test_custom_update_returns_1(lib_path):
    """acados_solver_bluerov2.c:1030-1036: prints its two lines and returns 1"""
    out = _a10("custom_update")
    assert out.returncode == 0 and "custom_update_rc 1" in out.stdout
    assert "dummy function that can be called in between solver calls" in out.stdout and "nothing set yet.." in out.stdout


@pytest.mark.gpu
def test_reset_and_sparse_params_behaviour(lib_path, oracle):
    """_reset zeroes the iterate (acados_solver_bluerov2.c:797-830) and the next solve is the cold solve from zeros;
    _update_params_sparse (:886-943) fed index by index equals dense _update_params; both against the oracle"""
    out = _a10("gpu")
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    kv = {}
    for line in out.stdout.splitlines():
        parts = line.split()
        if len(parts) >= 2 and parts[0].replace("_", "").isalnum():
            try:
                kv[parts[0]] = [float(v) for v in parts[1:]]
            except ValueError:
                pass
    assert kv["status_A"] == [0] and kv["status_after_reset"] == [0] and kv["status_B"] == [0]
    assert kv["reset_rc"] == [0] and kv["sparse_rc"] == [0] and kv["free_rc"] == [0] and kv["custom_update_rc"] == [1]
    assert kv["iterate_before_reset"][0] > 1.0 and kv["iterate_after_reset"] == [0.0]
    assert kv["reset_vs_cold"][0] == 0.0            # same engine, same inputs: bit-identical
    assert kv["reset_vs_warm"][0] > 1e-3            # and it is a different linearisation point than the warm start
    # the oracle from the zero iterate with the same data
    from bluerov2_b200 import workloads as wl
    N = 20
    x0 = np.zeros(12); x0[[0, 1, 2, 5, 6]] = [0.3, -0.2, -19.6, 0.25, 0.1]
    p = wl.NOMINAL_P.copy(); p[:4] = [1.5, -0.7, 0.4, 0.05]
    yref = np.zeros((N + 1, 16)); yref[:, 0] = 0.05 * np.arange(N + 1); yref[:, 2] = -20.0; yref[:, 6] = 1.0
    X, U = np.zeros((N + 1, 12)), np.zeros((N, 4))
    st, _ = oracle.rti_step(wl.time_steps(N), x0, yref, p, X, U)
    assert st == 0 and np.abs(U[0] - np.array(kv["u_after_reset"])).max() < 1e-6


# ---------------------------------------------------------------------------------------------------------
# the nine cost functions with CasADi's six entry points each (bluerov2_cost.h:45-116)
COST_FUNCS = [pre + suf for pre in ("bluerov2_cost_y_0", "bluerov2_cost_y", "bluerov2_cost_y_e") for suf in ("_fun", "_fun_jac_ut_xt", "_hess")]


def _casadi_api(lib, name, int_t):
    """evaluate `name` with random dense inputs sized by its own sparsity_in / sparsity_out, plus all helper results.
    int_t: the library's casadi_int (int on both sides: bluerov2_cost_y_fun.c:27)."""
    DP = C.POINTER(C.c_double)
    IP = C.POINTER(int_t)

    def sp(kind, i):
        fn = getattr(lib, f"{name}_sparsity_{kind}"); fn.argtypes = [int_t]; fn.restype = IP
        p = fn(i)
        if not p:
            return None
        rows, cols = p[0], p[1]
        colind = [p[2 + j] for j in range(cols + 1)]
        return (rows, cols, colind, [p[2 + cols + 1 + k] for k in range(colind[-1])])

    n = {}
    for k in ("n_in", "n_out"):
        fn = getattr(lib, f"{name}_{k}"); fn.argtypes = []; fn.restype = int_t
        n[k] = int(fn())
    sz = [int_t() for _ in range(4)]
    wk = getattr(lib, name + "_work"); wk.argtypes = [IP] * 4; wk.restype = C.c_int
    assert wk(*[C.byref(v) for v in sz]) == 0
    sp_in = [sp("in", i) for i in range(n["n_in"])]
    sp_out = [sp("out", i) for i in range(n["n_out"])]
    assert sp("in", n["n_in"]) is None and sp("out", n["n_out"]) is None
    rng = np.random.default_rng(abs(hash(name)) % 1000)
    ins = [rng.uniform(-2, 2, max(len(s[3]), 1)) for s in sp_in]
    outs = [np.full(max(len(s[3]), 1), np.nan) for s in sp_out]
    fn = getattr(lib, name); fn.argtypes = [C.POINTER(DP), C.POINTER(DP), C.c_void_p, C.c_void_p, C.c_void_p]; fn.restype = C.c_int
    arg = (DP * len(ins))(*[a.ctypes.data_as(DP) for a in ins])
    res = (DP * len(outs))(*[a.ctypes.data_as(DP) for a in outs])
    assert fn(arg, res, None, None, None) == 0
    vals = [o[:len(s[3])].copy() for o, s in zip(outs, sp_out)]
    # a null input reads as zeros
    arg0 = (DP * len(ins))(*[None for _ in ins])
    outs0 = [np.full(max(len(s[3]), 1), np.nan) for s in sp_out]
    res0 = (DP * len(outs0))(*[a.ctypes.data_as(DP) for a in outs0])
    assert fn(arg0, res0, None, None, None) == 0
    return dict(n=n, work=[int(v.value) for v in sz], sp_in=sp_in, sp_out=sp_out, ins=ins, vals=vals,
                vals_null=[o[:len(s[3])].copy() for o, s in zip(outs0, sp_out)])


@pytest.mark.parametrize("name", COST_FUNCS)
def test_cost_functions_known_answers(lib_path, name):
    """y = [x; u] (terminal y = x), Jacobian = one unit entry per column at the row of that variable in [u; x], empty Hessian"""
    r = _casadi_api(C.CDLL(lib_path), name, C.c_int)
    terminal = "_y_e_" in name
    ny = 12 if terminal else 16
    x, u = r["ins"][0], r["ins"][1]
    if name.endswith("_hess"):
        assert r["n"] == {"n_in": 5, "n_out": 1} and r["work"] == [5, 1, 0, 0]
        assert r["sp_out"][0] == (ny, ny, [0] * (ny + 1), []) and r["vals"][0].size == 0
        return
    want_y = x[:12] if terminal else np.concatenate([x[:12], u[:4]])
    assert np.array_equal(r["vals"][0], want_y) and np.array_equal(r["vals_null"][0], np.zeros(ny))
    if name.endswith("_fun"):
        assert r["n"] == {"n_in": 4, "n_out": 1} and r["work"] == [4, 1, 0, 0]
    else:
        assert r["n"] == {"n_in": 4, "n_out": 3} and r["work"] == [4, 3, 0, 0]
        rows = list(range(12)) if terminal else list(range(4, 16)) + [0, 1, 2, 3]
        assert r["sp_out"][1] == (ny, ny, list(range(ny + 1)), rows) and np.array_equal(r["vals"][1], np.ones(ny))
        assert r["sp_out"][2] == (ny, 0, [0], [])


@pytest.mark.parametrize("name", COST_FUNCS)
def test_cost_functions_match_reference_generated_c(lib_path, name):
    """all six entry points of every cost function against the reference's own CasADi-generated C (oracle/_ref, compiled
    from /root/reference/.../bluerov2_cost/*.c): n_in / n_out / work sizes, every sparsity pattern, and the values"""
    from oracle.oracle import REF_SO
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ours = _casadi_api(C.CDLL(lib_path), name, C.c_int)
    ref = _casadi_api(C.CDLL(REF_SO), name, C.c_int)            # casadi_int is int in the acados-generated code (bluerov2_cost_y_fun.c:27)
    assert ours["n"] == ref["n"] and ours["work"] == ref["work"]
    assert ours["sp_in"] == ref["sp_in"] and ours["sp_out"] == ref["sp_out"]
    for a, b in zip(ours["vals"], ref["vals"]):
        assert np.array_equal(a, b)
    for a, b in zip(ours["vals_null"], ref["vals_null"]):
        assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,needle", [("lbx_ne_ubx", "differs from ubx_0[3]"), ("scaling_ne_ts", "cost scaling 0.123 of stage 2 differs from its time step")])
def test_unrepresentable_settings_are_refused_loudly(lib_path, mode, needle):
    """lbx_0 != ubx_0 and cost scaling != time step are accepted by acados; this solver fixes x0 and ties scaling to Ts, so it
    says so and exits like it does for every other unsupported setting (no silent different problem)"""
    out = _a10(mode)
    assert out.returncode == 1, (out.returncode, out.stdout[-500:], out.stderr[-500:])
    assert needle in out.stderr and "returned" not in out.stdout
