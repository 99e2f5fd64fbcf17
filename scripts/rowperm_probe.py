"""Default path (no row permutation) against nkp_create_rowperm (LargeDiag, csrc/rowperm.cpp) on one operand:
fill, factor time, refinement steps, accuracy.  usage: python scripts/rowperm_probe.py [imt jmt km]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nk_ocn_tracer_jacobian_precond_b200 import solver, synth  # noqa: E402

shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (100, 116, 60)
g = synth.make_grid(*shape, seed=1)
c = synth.make_circulation(g, seed=1)
n, rp, ci, nz, (ii, jj, kk, _) = synth.assemble_crs(g, c)
import scipy.sparse as sp  # noqa: E402
A = sp.csr_matrix((nz, ci, rp), shape=(n, n))
rng = np.random.default_rng(0)
xs = rng.standard_normal((n, 8))
B = np.asfortranarray(A @ xs)
t = time.time()
info = solver.rowperm_largediag(n, rp, ci, nz)
t_match = time.time() - t
print(f"n={n} nnz={len(nz)} matching {t_match:.2f} s on the host, rows moved {(info[0] != np.arange(n)).sum()}", flush=True)
for name, rowperm in (("default (no row permutation)", None), ("LargeDiag row permutation + scalings", info)):
    t = time.time()
    s = solver.TracerJacobianSolver(n, rp, ci, coords=(ii, jj, kk), rowperm=rowperm)
    t_an = time.time() - t
    s.factor(nz)
    s.factor(nz)
    X = B.copy(order="F")
    berr = s.solve(X)
    st = s.stats()
    err = (np.linalg.norm(X - xs, axis=0) / np.linalg.norm(xs, axis=0)).max()
    res = (np.linalg.norm(A @ X - B, axis=0) / np.linalg.norm(B, axis=0)).max()
    print(f"{name}: analysis {t_an:.2f} s, factor {st['t_factor'] * 1e3:.1f} ms ({st['factor_flops']:.3e} flop, nnz(L+U) {st['nnz_lu']:.3e}), "
          f"solve 8 RHS {st['t_solve'] * 1e3:.2f} ms in {st['refine_steps']} steps, relres {res:.2e}, err vs x* {err:.2e}, "
          f"berr {berr.max():.1e}, tiny pivots {st['tiny_pivots']}", flush=True)
    s.close()
