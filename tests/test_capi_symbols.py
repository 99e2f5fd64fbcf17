"""The C-ABI libraries load and export every function include/*.h declares (no compute)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "nk_ocn_tracer_jacobian_precond_b200")


def _declared(header):
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"#define[^\n]*(\\\n[^\n]*)*", "", txt)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", txt)
    return sorted(set(n for n in names if n not in ("defined", "sizeof")))


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(PKG, "csrc")])


@pytest.mark.parametrize("header,lib", [
    ("include/nkprecond.h", "libnkprecond.so"),
    ("include/compat/superlu_ddefs.h", "libnkprecond.so"),
    ("include/compat/mpi.h", "libnkprecond.so"),
    ("include/compat/netcdf.h", "libnkp_nc3.so"),
    ("include/nkp_nc3.h", "libnkp_nc3.so"),
])
def test_library_exports_header_symbols(header, lib):
    path = os.path.join(PKG, lib)
    if not os.path.exists(path):
        _build()
    handle = ctypes.CDLL(path)
    names = _declared(os.path.join(ROOT, header))
    assert len(names) >= (1 if header.endswith("nkp_nc3.h") else 5)
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing


def test_python_binding_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device nkp_create must fail, not compute."""
    import numpy as np
    from nk_ocn_tracer_jacobian_precond_b200 import solver
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    rp = np.array([0, 1, 2], dtype=np.int32)
    ci = np.array([0, 1], dtype=np.int32)
    with pytest.raises(solver.NkpError):
        solver.TracerJacobianSolver(2, rp, ci)


def test_native_driver_usage_and_io_errors(tmp_path):
    """solve_ABbatch mirrors the reference's command-line behaviour (src/solve_ABglobal.c:38-99): usage
    errors and unreadable files exit with EXIT_FAILURE; no GPU is needed to get that far."""
    prog = os.path.join(PKG, "solve_ABbatch")
    if not os.path.exists(prog):
        _build()
    r = subprocess.run([prog, "-h"], capture_output=True, text=True)
    assert r.returncode == 1 and "usage: jacobian_precond [-D dbg_lvl] [-n nprow[,npcol]] [-v vars] matrix_fname inout_fname" in r.stderr
    r = subprocess.run([prog, "-D", "x", "a", "b"], capture_output=True, text=True)
    assert r.returncode == 1 and "error parsing argument 'x' for option 'D'" in r.stderr
    r = subprocess.run([prog, "only_one"], capture_output=True, text=True)
    assert r.returncode == 1 and "unexpected number of arguments" in r.stderr
    r = subprocess.run([prog, str(tmp_path / "missing.nc"), str(tmp_path / "t.nc")], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("(0) ")
