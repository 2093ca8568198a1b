"""In-tree build of the CUDA library (nvcc, sm_100a only).

    python -m bluerov2_b200.build [--force]
    python -m bluerov2_b200.build --stage DIR     # drop-in tree for the reference's CMake (see stage())

Produces, under ``bluerov2_b200/lib/``:

* ``libacados_ocp_solver_bluerov2.so`` -- the product: the batched engine (include/bluerov2_b200.h), the
  acados-generated solver ABI (include/acados_solver_bluerov2.h) and the slice of the acados C interface the
  reference's nodes call (include/acados_c/ocp_nlp_interface.h).  Same file name as the library the reference's
  CMake links (bluerov2_dobmpc/CMakeLists.txt:41,95).
* ``libacados.so``, ``libhpipm.so``, ``libblasfeo.so`` -- empty link shims that only carry a DT_NEEDED on the
  library above, so ``-lacados -lhpipm -lblasfeo`` link lines (c_generated_code/Makefile:98-110,
  CMakeLists.txt:96) resolve without the real libraries.

nvcc cross-compiles without a GPU; the built ``.so`` files are git-ignored but travel to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# BR2_VARIANT=<tag> builds / loads a tuning or instrumentation variant from lib_<tag>/ (with BR2_NVCC_DEFS), leaving lib/ alone
_VARIANT = os.environ.get("BR2_VARIANT", "")
LIBDIR = os.path.join(HERE, "lib_" + _VARIANT if _VARIANT else "lib")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(LIBDIR, "libacados_ocp_solver_bluerov2.so")
SHIMS = ("libacados.so", "libhpipm.so", "libblasfeo.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps() -> list[str]:
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    for r, _, fs in os.walk(INCLUDE):
        d += [os.path.join(r, f) for f in fs]
    return d


def up_to_date() -> bool:
    if not os.path.exists(LIB) or not all(os.path.exists(os.path.join(LIBDIR, s)) for s in SHIMS):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(p) <= t for p in _deps())


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    for src in sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        extra = os.environ.get("BR2_NVCC_DEFS", "").split()     # tuning experiments, e.g. "-DBR2_NSLOT=3 -DBR2_IPM_MINB=5"
        cmd = [nvcc, *ARCH, *NVCC_FLAGS, *extra, "-Xptxas", "-v",
               "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
        out = subprocess.run(cmd, capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out.stdout + out.stderr)
        if verbose:
            print(out.stderr)
        with open(obj[:-2] + ".ptxas.txt", "w") as f:
            f.write(out.stderr)
        objs.append(obj)
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-Xlinker", "-soname,libacados_ocp_solver_bluerov2.so",
           "-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + out.stdout + out.stderr)
    # link shims: empty libraries whose only content is the dependency on the product library
    stub = os.path.join(LIBDIR, "_stub.c")
    with open(stub, "w") as f:
        f.write("/* link shim: see bluerov2_b200/build.py */\n")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    for s in SHIMS:
        cmd = [cc, "-shared", "-fPIC", "-o", os.path.join(LIBDIR, s), stub, "-Wl,--no-as-needed", "-L" + LIBDIR,
               "-lacados_ocp_solver_bluerov2", "-Wl,-rpath,$ORIGIN", "-Wl,-soname," + s]
        out = subprocess.run(cmd, capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError("shim link failed:\n" + out.stdout + out.stderr)
    os.remove(stub)
    return LIB


def stage(dest: str) -> str:
    """Assemble the drop-in tree the reference's CMake expects (bluerov2_dobmpc/CMakeLists.txt:39-41,71-78,94-97):

        <dest>/c_generated_code/   stands in for bluerov2_dobmpc/scripts/c_generated_code (``${bluerov2_model}``):
                                   acados_solver_bluerov2.h, bluerov2_model/, bluerov2_cost/, bluerov2_constraints/,
                                   libacados_ocp_solver_bluerov2.so
        <dest>/acados/include/     stands in for ~/acados/include (``${acados_include}``): acados/, acados_c/, blasfeo/
        <dest>/acados/lib/         stands in for ~/acados/lib (``${acados_lib}``): libacados.so, libhpipm.so, libblasfeo.so
        <dest>/cmake/              bluerov2_b200-config.cmake (find_package(bluerov2_b200 CONFIG))

    With it the reference's CMakeLists needs exactly the two acados paths it already asks every user to set; the
    generated directory is replaced wholesale.  The shims find the product library through their $ORIGIN rpath
    (../../c_generated_code)."""
    build_library()
    dest = os.path.abspath(dest)
    gen, ainc, alib = (os.path.join(dest, "c_generated_code"), os.path.join(dest, "acados", "include"),
                       os.path.join(dest, "acados", "lib"))
    for d in (gen, ainc, alib, os.path.join(dest, "cmake")):
        os.makedirs(d, exist_ok=True)
    shutil.copy2(os.path.join(INCLUDE, "acados_solver_bluerov2.h"), gen)
    shutil.copy2(os.path.join(INCLUDE, "bluerov2_b200.h"), gen)
    for sub in ("bluerov2_model", "bluerov2_cost", "bluerov2_constraints"):
        shutil.copytree(os.path.join(INCLUDE, sub), os.path.join(gen, sub), dirs_exist_ok=True)
    for sub in ("acados", "acados_c", "blasfeo"):
        shutil.copytree(os.path.join(INCLUDE, sub), os.path.join(ainc, sub), dirs_exist_ok=True)
    shutil.copy2(LIB, gen)
    stub = os.path.join(alib, "_stub.c")
    with open(stub, "w") as f:
        f.write("/* link shim: see bluerov2_b200/build.py */\n")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    for sh in SHIMS:
        cmd = [cc, "-shared", "-fPIC", "-o", os.path.join(alib, sh), stub, "-Wl,--no-as-needed", "-L" + gen,
               "-lacados_ocp_solver_bluerov2", "-Wl,-rpath,$ORIGIN/../../c_generated_code", "-Wl,-soname," + sh]
        out = subprocess.run(cmd, capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError("shim link failed:\n" + out.stdout + out.stderr)
    os.remove(stub)
    shutil.copy2(os.path.join(ROOT, "cmake", "bluerov2_b200-config.cmake"), os.path.join(dest, "cmake"))
    return dest


if __name__ == "__main__":
    if "--stage" in sys.argv:
        print(stage(sys.argv[sys.argv.index("--stage") + 1]))
    else:
        p = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
        print(p)
