"""Assembly tree A/B on one operand: NKP_SUPERNODES=0 (one dense front per dissection node, the round-1 tree) against
NKP_SUPERNODES=1 (relaxed supernodes of the elimination tree, csrc/analysis.cpp::supernodes_from_etree).
usage: python scripts/tree_ab.py imt jmt km [variants, e.g. 0,1]   (variant "1:0.2:64" = supernodes with relax_frac 0.2, relax_small 64)"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nk_ocn_tracer_jacobian_precond_b200 import solver, synth  # noqa: E402

shape = tuple(int(v) for v in sys.argv[1:4])
variants = sys.argv[4].split(",") if len(sys.argv) > 4 else ["0", "1"]
t = time.time()
g = synth.make_grid(*shape, seed=1)
c = synth.make_circulation(g, seed=1)
n, rp, ci, nz, (ii, jj, kk, _) = synth.assemble_crs(g, c)
A = sp.csr_matrix((nz, ci, rp), shape=(n, n))
rng = np.random.default_rng(0)
xs = rng.standard_normal((n, 8))
B = np.asfortranarray(A @ xs)
print(f"n={n} nnz={len(nz)} (operand built in {time.time() - t:.1f} s)", flush=True)
for v in variants:
    parts = v.split(":")
    os.environ["NKP_SUPERNODES"] = parts[0]
    if len(parts) > 1:
        os.environ["NKP_RELAX_FRAC"] = parts[1]
    if len(parts) > 2:
        os.environ["NKP_RELAX_SMALL"] = parts[2]
    t = time.time()
    s = solver.TracerJacobianSolver(n, rp, ci, coords=(ii, jj, kk))
    t_an = time.time() - t
    tf = []
    for _ in range(3):
        s.factor(nz)
        tf.append(s.stats()["t_factor"])
    ts, steps = [], 0
    for _ in range(2):
        X = B.copy(order="F")
        berr = s.solve(X)
        st = s.stats()
        ts.append(st["t_solve"])
        steps = st["refine_steps"]
    err = (np.linalg.norm(X - xs, axis=0) / np.linalg.norm(xs, axis=0)).max()
    res = (np.linalg.norm(A @ X - B, axis=0) / np.linalg.norm(B, axis=0)).max()
    print(f"variant {v}: analysis {t_an:.2f} s, fronts {st['n_fronts']}, levels {st['n_levels']}, max front {st['max_front']}, "
          f"flops {st['factor_flops']:.4g}, nnz(L+U) {st['nnz_lu']:.4g}, heap {st['heap_bytes'] / 1e9:.2f} GB\n"
          f"   factor {min(tf) * 1e3:.2f} ms ({st['factor_flops'] / min(tf) * 1e-12:.2f} TF/s on the plan's flops), "
          f"solve 8 RHS {min(ts) * 1e3:.2f} ms in {steps} steps, launches {st['kernel_launches']}, "
          f"relres {res:.2e}, err vs x* {err:.2e}, berr {berr.max():.1e}, tiny pivots {st['tiny_pivots']}", flush=True)
    s.close()
