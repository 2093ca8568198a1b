"""ROS-side packaging: the staged drop-in tree + the reference's CMake statements configure, build
and link with nothing but the two acados paths changed.  CPU-only: proves the link line; running needs the GPU box."""
import os
import shutil
import subprocess

import pytest

from tests.conftest import REFERENCE, ROOT

cmake = shutil.which("cmake")
pytestmark = pytest.mark.skipif(cmake is None, reason="cmake not available")


@pytest.fixture(scope="module")
def staged(tmp_path_factory):
    from bluerov2_b200 import build
    d = tmp_path_factory.mktemp("stage")
    build.stage(str(d))
    return str(d)


def test_staged_tree_layout(staged):
    for rel in ("c_generated_code/acados_solver_bluerov2.h", "c_generated_code/bluerov2_model/bluerov2_model.h",
                "c_generated_code/libacados_ocp_solver_bluerov2.so", "acados/include/acados_c/ocp_nlp_interface.h",
                "acados/include/blasfeo/include/blasfeo_d_aux.h", "acados/lib/libacados.so", "acados/lib/libhpipm.so",
                "acados/lib/libblasfeo.so", "cmake/bluerov2_b200-config.cmake"):
        assert os.path.exists(os.path.join(staged, rel)), rel
    # the shim resolves the product library relative to itself
    out = subprocess.run(["ldd", os.path.join(staged, "acados/lib/libacados.so")], capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if "libacados_ocp_solver_bluerov2.so" in l]
    assert line and "not found" not in line[0] and os.path.realpath(staged) in os.path.realpath(line[0].split("=>")[1].split()[0])


@pytest.mark.parametrize("use_package", [False, True])
def test_reference_cmake_statements_build_against_staged_tree(staged, tmp_path, use_package):
    bdir = tmp_path / "build"
    cfg = [cmake, "-S", os.path.join(ROOT, "tests", "cmake_dropin"), "-B", str(bdir), f"-DBR2_STAGE={staged}",
           f"-DBR2_REPO={ROOT}", f"-DBR2_REFERENCE={REFERENCE}", f"-DBR2_USE_PACKAGE={'ON' if use_package else 'OFF'}",
           "-DCMAKE_C_COMPILER=/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "-DCMAKE_C_COMPILER=gcc"]
    out = subprocess.run(cfg, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    out = subprocess.run([cmake, "--build", str(bdir)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    exes = ["dob_tick"] + (["main_bluerov2"] if os.path.isdir(REFERENCE) else [])
    for exe in exes:
        path = bdir / exe
        assert path.exists(), exe
        ldd = subprocess.run(["ldd", str(path)], capture_output=True, text=True).stdout
        line = [l for l in ldd.splitlines() if "libacados_ocp_solver_bluerov2.so" in l]
        assert line and "not found" not in line[0], (exe, ldd)
        # the libacados.so shim defines nothing, so a toolchain linking with --as-needed drops it; if it is kept it must resolve
        assert not [l for l in ldd.splitlines() if "not found" in l], (exe, ldd)
