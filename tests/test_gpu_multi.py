"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): one process per GPU,
NCCL inside the solver, answers identical to the single-GPU solve (KAT-6)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, numpy as np
sys.path.insert(0, os.environ["NKP_ROOT"])
import torch, torch.distributed as dist
from nk_ocn_tracer_jacobian_precond_b200 import solver, synth
import scipy.sparse as sp
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = synth.make_grid(40, 46, 24, seed=1); c = synth.make_circulation(g, seed=1)
n, rp, ci, nz, (ii, jj, kk, _) = synth.assemble_crs(g, c)
uid = [solver.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
s = solver.TracerJacobianSolver(n, rp, ci, coords=(ii, jj, kk), comm=(rank, world, uid[0]), device=local)
A = sp.csr_matrix((nz, ci, rp), shape=(n, n))
xs = np.random.default_rng(0).standard_normal((n, 8))
B = np.asfortranarray(A @ xs)
for rep in range(2):
    s.factor(nz)
    X = B.copy(order="F"); berr = s.solve(X)
res = (np.linalg.norm(A @ X - B, axis=0) / np.linalg.norm(B, axis=0)).max()
err = (np.linalg.norm(X - xs, axis=0) / np.linalg.norm(xs, axis=0)).max()
st = s.stats()
s.close()
ok = res <= 1e-10 and err <= 1e-8 and st["factor_flops_local"] < 0.9 * st["factor_flops"] and st["n_xfers"] > 0
if rank == 0:
    # single-GPU answer for comparison
    s1 = solver.TracerJacobianSolver(n, rp, ci, coords=(ii, jj, kk), device=local)
    s1.factor(nz); X1 = B.copy(order="F"); s1.solve(X1); s1.close()
    diff = np.linalg.norm(X1 - X) / np.linalg.norm(X1)
    print(f"MULTI world={world} res={res:.2e} err={err:.2e} diff_vs_1gpu={diff:.2e} local_flops_frac={st['factor_flops_local']/st['factor_flops']:.2f}")
    ok = ok and diff <= 1e-10
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("world,outer", [(2, 8), (2, 1), (4, 8), (4, 1), (8, 2)])
def test_multi_gpu_matches_single_gpu(tmp_path, world, outer):
    """outer (NKP_OUTER) sets the width of the distribution blocks of the top fronts (outer * 64 columns): with 1 or 2
    the 40x46x24 grid gives every top front many blocks, so owner rotation, look-ahead panels, panel broadcasts and
    the column-block broadcasts of distributed update matrices are all exercised."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, NKP_ROOT=ROOT, NKP_OUTER=str(outer))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29600 + 10 * world + outer), str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    print(out.stdout[-400:])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MULTI" in out.stdout
