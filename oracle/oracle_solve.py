"""oracle_solve.py -- TEST INFRASTRUCTURE: CPU oracle of the reference's solve path.

The reference delegates the arithmetic of this path to SuperLU_DIST 5.1.3
(/root/reference/src/Makefile:3; call sites src/solve_ABglobal.c:353,395 and
src/solve_ABdist.c:518,571), which is NOT in the reference tree and cannot be built in
this image (no MPI / ParMETIS / BLAS).  The oracle therefore restates
the published algorithm of pdgssvx with the same library family: scipy's bundled *serial*
SuperLU (``scipy.sparse.linalg.splu``: supernodal LU with partial pivoting) followed by
the iterative refinement loop of pdgsrfs (src/SuperLU_brief_tree.txt:20-24: residual,
|A||x|+|b|, correction solve; stop when berr <= eps, when berr fails to halve, or after
ITMAX steps).

PARITY UNPINNED BY THE REFERENCE: the reference holds no golden vectors, fixtures or
tolerances for this path.  What pins this oracle instead:
  * the operand: CRS produced by the reference's own, unchanged gen_A (oracle/_ref/gen_A,
    golden file tests/golden/A_20x24x10.nc) is what the oracle and the GPU path both read;
  * self-consistency: residual of the oracle's own solution (tests/test_oracle.py).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

EPS = np.finfo(np.float64).eps
ITMAX = 20


def csr(n, rowptr, colind, nzval):
    return sp.csr_matrix((np.asarray(nzval, dtype=np.float64), np.asarray(colind), np.asarray(rowptr)), shape=(n, n))


def spmv(n, rowptr, colind, nzval, x):
    return csr(n, rowptr, colind, nzval) @ x


def factor(n, rowptr, colind, nzval, permc_spec="COLAMD"):
    """pdgssvx with nrhs = 0 (src/solve_ABglobal.c:353): returns the LU object."""
    return spla.splu(csr(n, rowptr, colind, nzval).tocsc(), permc_spec=permc_spec)


def berr_of(A, x, b):
    """Componentwise backward error max_i |r_i| / (|A||x|+|b|)_i (pdgsrfs)."""
    r = b - A @ x
    den = abs(A) @ np.abs(x) + np.abs(b)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.where(den > 0, np.abs(r) / den, 0.0)
    return float(q.max()), r


def refine(A, lu, b, x):
    """pdgsrfs stopping rule (src/SuperLU_brief_tree.txt:20-24)."""
    last = np.inf
    steps = 0
    while True:
        berr, r = berr_of(A, x, b)
        if not (berr > EPS and berr * 2.0 <= last and steps < ITMAX):
            return x, berr, steps
        last = berr
        x = x + lu.solve(r)
        steps += 1


def solve(n, rowptr, colind, nzval, B, lu=None, return_info=False):
    """pdgssvx with Fact = FACTORED, nrhs >= 1 (src/solve_ABglobal.c:395): X = A^-1 B."""
    A = csr(n, rowptr, colind, nzval)
    if lu is None:
        lu = spla.splu(A.tocsc())
    B = np.asarray(B, dtype=np.float64)
    cols = B.reshape(n, -1, order="F")
    X = np.empty_like(cols)
    info = []
    for c in range(cols.shape[1]):
        x0 = lu.solve(cols[:, c])
        x, berr, steps = refine(A, lu, cols[:, c], x0)
        X[:, c] = x
        info.append((berr, steps))
    X = X.reshape(B.shape, order="F")
    return (X, info) if return_info else X
