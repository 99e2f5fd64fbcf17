"""Per-level timing of one sweep pair (NKP_VERBOSE=3)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nk_ocn_tracer_jacobian_precond_b200 import solver
wl = sys.argv[1] if len(sys.argv) > 1 else "gx3v7"
nrhs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
case = bench.build_case(wl)
s = solver.TracerJacobianSolver(case["n"], case["rowptr"], case["colind"], coords=case["coords"], verbose=0)
s.factor(case["nzval"])
B = torch.randn(nrhs, case["n"], dtype=torch.float64, device="cuda")
for it in range(2):
    s.sweeps_device(B.data_ptr(), case["n"], nrhs)
print("t_sweeps ms", s.stats()["t_sweeps"] * 1e3)
s2 = None
os.environ["NKP_VERBOSE"] = "3"
s.close()
s = solver.TracerJacobianSolver(case["n"], case["rowptr"], case["colind"], coords=case["coords"])
s.factor(case["nzval"])
s.sweeps_device(B.data_ptr(), case["n"], nrhs)
s.sweeps_device(B.data_ptr(), case["n"], nrhs)
