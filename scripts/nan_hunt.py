import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from nk_ocn_tracer_jacobian_precond_b200 import solver
case = bench.build_case(sys.argv[1] if len(sys.argv) > 1 else "gx1v6")
n, nnz = case["n"], len(case["nzval"])
s = solver.TracerJacobianSolver(n, case["rowptr"], case["colind"], coords=case["coords"])
rng = np.random.default_rng(1234)
vals = [torch.tensor(case["nzval"] * (1.0 + 1e-3 * rng.standard_normal(nnz)), device="cuda") for _ in range(3)]
b = torch.tensor(rng.standard_normal((8, n)), device="cuda")
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 8):
    s.factor_device(vals[it % 3].data_ptr())
    w = b.clone()
    berr = s.solve_device(w.data_ptr(), n, 8)
    print(f"it {it}: steps {s.stats()['refine_steps']} berr {berr.max():.2e} finite {bool(torch.isfinite(w).all())}", flush=True)
s.close()
