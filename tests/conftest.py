import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_matrix():
    from nk_ocn_tracer_jacobian_precond_b200 import synth
    m = synth.read_matrix_file(os.path.join(GOLDEN, "A_20x24x10.nc"))
    m["n"] = len(m["rowptr"]) - 1
    return m


@pytest.fixture(scope="session")
def golden_rhs():
    import numpy as np
    d = np.load(os.path.join(GOLDEN, "rhs_x_20x24x10.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def reftest_matrix():
    """Operand of the reference's own test option set (upwind3 + isop_file + vmix file)."""
    from nk_ocn_tracer_jacobian_precond_b200 import synth
    m = synth.read_matrix_file(os.path.join(GOLDEN, "A_reftest_20x24x10.nc"))
    m["n"] = len(m["rowptr"]) - 1
    return m


@pytest.fixture(scope="session")
def reftest_rhs():
    import numpy as np
    d = np.load(os.path.join(GOLDEN, "rhs_x_reftest_20x24x10.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def sim_lib():
    """CPU plan interpreter (oracle/libnkp_sim.so); built on demand."""
    import ctypes
    import subprocess
    path = os.path.join(ROOT, "oracle", "libnkp_sim.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "sim"])
    return ctypes.CDLL(path)


def synth_case(imt, jmt, km, seed=1):
    from nk_ocn_tracer_jacobian_precond_b200 import synth
    g = synth.make_grid(imt, jmt, km, seed=seed)
    c = synth.make_circulation(g, seed=seed)
    n, rp, ci, nz, (ii, jj, kk, int3) = synth.assemble_crs(g, c)
    return dict(grid=g, circ=c, n=n, rowptr=rp, colind=ci, nzval=nz, i=ii, j=jj, k=kk, int3=int3)
