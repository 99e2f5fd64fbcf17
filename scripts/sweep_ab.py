"""A/B of the sweep kernels on one GPU: sweep-pair time, per-launch trace (verbose 3), parity."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import scipy.sparse as sp
import bench
from nk_ocn_tracer_jacobian_precond_b200 import solver

wl = sys.argv[1] if len(sys.argv) > 1 else "gx3v7"
trace = int(sys.argv[2]) if len(sys.argv) > 2 else 0
case = bench.build_case(wl)
n = case["n"]
A = sp.csr_matrix((case["nzval"], case["colind"], case["rowptr"]), shape=(n, n))
xs = np.random.default_rng(0).standard_normal((n, 8))
B = bench.spmv_extended(case["rowptr"], case["colind"], case["nzval"], xs)
s = solver.TracerJacobianSolver(n, case["rowptr"], case["colind"], coords=case["coords"])
s.factor(case["nzval"])
db = torch.tensor(np.ascontiguousarray(B.T), device="cuda")
ts = []
for it in range(6):
    w = db.clone()
    if it == 5 and trace:
        s._lib.nkp_set_verbose(s._h, 3)
    s.sweeps_device(w.data_ptr(), n, 8)
    s._lib.nkp_set_verbose(s._h, 0)
    ts.append(s.stats()["t_sweeps"] * 1e3)
st = s.stats()
print(f"{wl} small_v1={os.environ.get('NKP_SMALL_V1', '0')}: sweep pair ms {min(ts[1:5]):.3f}  ({st['solve_bytes'] / min(ts[1:5]) * 1e-6:.0f} GB/s)")
for nr in (8, 3, 1):
    w = db[:nr].clone().contiguous()
    berr = s.solve_device(w.data_ptr(), n, nr)
    X = w.cpu().numpy().T
    res = (np.linalg.norm(A @ X - B[:, :nr], axis=0) / np.linalg.norm(B[:, :nr], axis=0)).max()
    err = (np.linalg.norm(X - xs[:, :nr], axis=0) / np.linalg.norm(xs[:, :nr], axis=0)).max()
    st = s.stats()
    print(f"  solve nrhs={nr}: {st['t_solve'] * 1e3:.2f} ms, steps {st['refine_steps']}, berr {berr.max():.2e}, relres {res:.2e}, err {err:.2e}")
s.close()
