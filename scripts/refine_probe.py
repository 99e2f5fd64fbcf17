"""GPU probe: convergence history of iterative refinement (componentwise berr, normwise residual and error
against a manufactured solution, per step) with the working-precision and the extra-precise residual.

  python scripts/refine_probe.py gx1v6 [nrhs]
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scipy.sparse as sp
import torch
import bench
from nk_ocn_tracer_jacobian_precond_b200 import solver

wl = sys.argv[1] if len(sys.argv) > 1 else "gx1v6"
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 2
case = bench.build_case(wl)
n = case["n"]
A = sp.csr_matrix((case["nzval"], case["colind"], case["rowptr"]), shape=(n, n))
absA = abs(A)
rng = np.random.default_rng(0)
xs = rng.standard_normal((n, nr))
b_ext = bench.spmv_extended(case["rowptr"], case["colind"], case["nzval"], xs)
b_dbl = np.asfortranarray(A @ xs)
print(f"{wl}: n={n} nnz={A.nnz}; |b_dbl - b_ext| / |b| max = {np.abs(b_dbl - b_ext).max() / np.abs(b_ext).max():.2e}", flush=True)
s = solver.TracerJacobianSolver(n, case["rowptr"], case["colind"], coords=case["coords"])
s.factor(case["nzval"])
st = s.stats()
print(f"factor {st['t_factor']:.3f} s, tiny pivots {st['tiny_pivots']}", flush=True)
dev = torch.device("cuda", 0)
for bname, b in (("b double", b_dbl), ("b extended", b_ext)):
    for extra in (0, 1):
        s.set_residual_extra(extra)
        db = torch.tensor(np.ascontiguousarray(b.T), device=dev)
        x = torch.zeros_like(db)
        r = db.clone()
        print(f"--- {bname}, residual_extra={extra}")
        for it in range(10):
            d = r.clone()
            s.sweeps_device(d.data_ptr(), n, nr)
            x += d
            s.residual_device(x.data_ptr(), db.data_ptr(), r.data_ptr(), nr)
            s.sync()
            X = x.cpu().numpy().T
            R = r.cpu().numpy().T
            den = absA @ np.abs(X) + np.abs(b)
            berr = (np.abs(R) / den).max(axis=0)
            relres = np.linalg.norm(R, axis=0) / np.linalg.norm(b, axis=0)
            err = np.linalg.norm(X - xs, axis=0) / np.linalg.norm(xs, axis=0)
            errinf = np.abs(X - xs).max(axis=0) / np.abs(xs).max(axis=0)
            print(f"  it {it}: berr {berr.max():.2e} relres {relres.max():.2e} err2 {err.max():.3e} errinf {errinf.max():.3e}", flush=True)
        # the library's own loop
        for rule in (0, 1):
            s.set_refine_rule(rule)
            w = db.clone()
            berr = s.solve_device(w.data_ptr(), n, nr)
            X = w.cpu().numpy().T
            err = np.linalg.norm(X - xs, axis=0) / np.linalg.norm(xs, axis=0)
            sst = s.stats()
            print(f"  nkp_solve rule {rule}: steps {sst['refine_steps']} t {sst['t_solve'] * 1e3:.1f} ms berr {berr.max():.2e} err2 {err.max():.3e}", flush=True)
        s.set_refine_rule(0)
s.close()
