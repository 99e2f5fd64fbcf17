"""Developer check on a GPU box: correctness + timing of factor/solve at several sizes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nk_ocn_tracer_jacobian_precond_b200 import solver, synth
import scipy.sparse as sp

def run(imt, jmt, km, nrhs=1, reps=int(os.environ.get("NKP_CHECK_REPS", "3"))):
    g = synth.make_grid(imt, jmt, km, seed=1); c = synth.make_circulation(g, seed=1)
    n, rp, ci, nz, (ii, jj, kk, _) = synth.assemble_crs(g, c)
    A = sp.csr_matrix((nz, ci, rp), shape=(n, n))
    t = time.time()
    s = solver.TracerJacobianSolver(n, rp, ci, coords=(ii, jj, kk), verbose=1)
    print(f"[{imt}x{jmt}x{km}] n={n} nnz={len(nz)} create {time.time()-t:.2f}s", flush=True)
    for r in range(reps):
        t = time.time(); s.factor(nz); tw = time.time() - t
        st = s.stats()
        print(f"  factor: wall {tw*1e3:.1f} ms dev {st['t_factor']*1e3:.2f} ms scatter {st['t_scatter']*1e3:.2f} ms "
              f"{st['factor_flops']/st['t_factor']*1e-12:.2f} TF/s tiny={st['tiny_pivots']}", flush=True)
    rng = np.random.default_rng(0)
    xs = rng.standard_normal((n, nrhs))
    B = np.asfortranarray(A @ xs)
    X = B.copy(order='F')
    for r in range(reps):
        X[:] = B
        t = time.time(); berr = s.solve(X); tw = time.time() - t
        st = s.stats()
        res = np.linalg.norm(A @ X - B, axis=0) / np.linalg.norm(B, axis=0)
        err = np.linalg.norm(X - xs, axis=0) / np.linalg.norm(xs, axis=0)
        print(f"  solve nrhs={nrhs}: wall {tw*1e3:.1f} ms dev {st['t_solve']*1e3:.2f} ms refine={st['refine_steps']} "
              f"relres max {res.max():.2e} err max {err.max():.2e} berr max {berr.max():.1e}", flush=True)
    b2 = np.asfortranarray(rng.standard_normal((n, 1)))
    x2 = b2.copy(order='F'); berr = s.solve(x2)
    print(f"  random rhs: relres {np.linalg.norm(A@x2-b2)/np.linalg.norm(b2):.2e} berr {berr[0]:.1e} refine={s.stats()['refine_steps']}", flush=True)
    print("  stats", {k: v for k, v in s.stats().items() if k in ('n_fronts','n_levels','max_front','nnz_lu','heap_bytes','kernel_launches','t_analysis')}, flush=True)
    s.close()

if __name__ == "__main__":
    sizes = sys.argv[1:] or ["20x24x10", "50x58x30", "100x116x60"]
    for sz in sizes:
        a = sz.split(":")
        imt, jmt, km = (int(v) for v in a[0].split("x"))
        run(imt, jmt, km, nrhs=int(a[1]) if len(a) > 1 else 1)
