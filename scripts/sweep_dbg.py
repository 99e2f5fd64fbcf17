import os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, bench
    from nk_ocn_tracer_jacobian_precond_b200 import solver
    case = bench.build_case(sys.argv[2])
    n = case["n"]
    s = solver.TracerJacobianSolver(n, case["rowptr"], case["colind"], coords=case["coords"], leaf=int(os.environ.get("LEAF", "96")))
    s.factor(case["nzval"])
    b = np.random.default_rng(0).standard_normal((8, n))
    for nr in (8, 1):
        w = torch.tensor(b[:nr].copy(), device="cuda")
        s.sweeps_device(w.data_ptr(), n, nr)
        np.save(sys.argv[3] + f"_{nr}.npy", w.cpu().numpy())
    st = s.stats()
    print("fronts", st["n_fronts"], "levels", st["n_levels"])
    s.close()
else:
    wl = sys.argv[1]
    for v in (3, 2, 1, 0):
        env = dict(os.environ, NKP_SMALL_V1=str(v))
        subprocess.check_call([sys.executable, __file__, "child", wl, f"/tmp/sw_{v}"], env=env)
    for nr in (8, 1):
        ref = np.load(f"/tmp/sw_3_{nr}.npy")
        for v in (2, 1, 0):
            x = np.load(f"/tmp/sw_{v}_{nr}.npy")
            d = np.abs(x - ref).max(axis=1) / np.abs(ref).max(axis=1)
            bad = np.argwhere(np.abs(x - ref) > 1e-6 * np.abs(ref).max())
            print(f"nr={nr} small_v1={v} (bit0: old fwd, bit1: old bwd): max rel diff per rhs {d.max():.2e}; wrong entries {len(bad)} of {x.size}", flush=True)
