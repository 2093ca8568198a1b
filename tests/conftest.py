import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle, build
    build()
    return Oracle()


@pytest.fixture(scope="session")
def casadi_ref():
    from oracle import CasadiRef
    from oracle.oracle import REF_SO as path
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return CasadiRef()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return {n: np.load(os.path.join(GOLDEN, n + ".npz")) for n in ("casadi_vde", "rti_cases", "ekf_cases")}
