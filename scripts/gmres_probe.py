"""Developer probe: stationary refinement vs right-preconditioned GMRES with the same LU factors.
Uses the library's raw sweep pair (nkp_sweeps_device = M^-1) and residual SpMV (nkp_residual_device)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nk_ocn_tracer_jacobian_precond_b200 import solver

wl = sys.argv[1] if len(sys.argv) > 1 else "gx3v7"
case = bench.build_case(wl)
n = case["n"]
s = solver.TracerJacobianSolver(n, case["rowptr"], case["colind"], coords=case["coords"])
s.factor(case["nzval"])
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
zero = torch.zeros(n, dtype=torch.float64, device=dev)


def A_mul(z):
    r = torch.empty_like(z)
    s.residual_device(z.data_ptr(), zero.data_ptr(), r.data_ptr(), 1)   # r = 0 - A z
    s.sync()
    return -r


def Minv(v):
    y = v.clone()
    s.sweeps_device(y.data_ptr(), n, 1)
    return y


for trial, kind in enumerate(["manufactured", "random"]):
    if kind == "manufactured":
        xs = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
        b = A_mul(xs)
    else:
        b = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    bn = b.norm().item()
    # stationary refinement
    x = Minv(b)
    hist = []
    for it in range(9):
        r = b - A_mul(x)
        hist.append(r.norm().item() / bn)
        x = x + Minv(r)
    print(kind, "refinement relres:", " ".join(f"{h:.1e}" for h in hist), flush=True)
    # right-preconditioned GMRES (modified Gram-Schmidt), one cycle
    m = 8
    V = [b / bn]
    Z = []
    H = np.zeros((m + 1, m))
    ghist = []
    for j in range(m):
        z = Minv(V[j])
        Z.append(z)
        w = A_mul(z)
        for i in range(j + 1):
            H[i, j] = torch.dot(w, V[i]).item()
            w = w - H[i, j] * V[i]
        H[j + 1, j] = w.norm().item()
        V.append(w / H[j + 1, j])
        e1 = np.zeros(j + 2); e1[0] = bn
        yk, *_ = np.linalg.lstsq(H[: j + 2, : j + 1], e1, rcond=None)
        xk = sum(float(yk[i]) * Z[i] for i in range(j + 1))
        ghist.append((b - A_mul(xk)).norm().item() / bn)
    print(kind, "GMRES      relres:", " ".join(f"{h:.1e}" for h in ghist), flush=True)
s.close()
