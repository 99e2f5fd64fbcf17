"""Multi-GPU developer probe (run under torchrun): factor timeline per rank (NKP verbose 2/3), sweep timings, parity.

  python -m torch.distributed.run --nproc-per-node N ... scripts/dist_trace.py gx1v6 [verbose]
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import scipy.sparse as sp
import bench
from nk_ocn_tracer_jacobian_precond_b200 import solver

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
comm = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = sys.argv[1] if len(sys.argv) > 1 else "gx1v6"
verbose = int(sys.argv[2]) if len(sys.argv) > 2 else 2
case = bench.build_case(wl)
n = case["n"]
if world > 1:
    uid = [solver.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = (rank, world, uid[0])
s = solver.TracerJacobianSolver(n, case["rowptr"], case["colind"], coords=case["coords"], comm=comm, device=local, verbose=0)
st = s.stats()
print(f"rank {rank}: local flops {st['factor_flops_local']:.3e} of {st['factor_flops']:.3e}, local nnz_lu {st['nnz_lu_local']:.3e}, heap {st['heap_bytes']*1e-9:.1f} GB", flush=True)
dv = torch.tensor(case["nzval"], device=f"cuda:{local}")
import ctypes
for it in range(3):
    if it == 2:
        # switch the timeline on for the last repetition only
        os.environ["NKP_TRACE"] = "1"
    if world > 1:
        dist.barrier()
    s._lib.nkp_set_verbose(s._h, verbose if it == 2 else 0)
    s.factor_device(dv.data_ptr())
    print(f"rank {rank}: factor {s.stats()['t_factor']*1e3:.1f} ms", flush=True)
s._lib.nkp_set_verbose(s._h, 0)
A = sp.csr_matrix((case["nzval"], case["colind"], case["rowptr"]), shape=(n, n))
xs = np.random.default_rng(0).standard_normal((n, 8))
B = bench.spmv_extended(case["rowptr"], case["colind"], case["nzval"], xs)
db = torch.tensor(np.ascontiguousarray(B.T), device=f"cuda:{local}")
for it in range(3):
    w = db.clone()
    if world > 1:
        dist.barrier()
    if it == 2 and verbose >= 3:
        s._lib.nkp_set_verbose(s._h, 3)
    berr = s.solve_device(w.data_ptr(), n, 8)
    s._lib.nkp_set_verbose(s._h, 0)
    stt = s.stats()
    if rank == 0:
        print(f"solve: {stt['t_solve']*1e3:.2f} ms, steps {stt['refine_steps']}, berr {berr.max():.2e}", flush=True)
for it in range(3):
    w = db.clone()
    if world > 1:
        dist.barrier()
    s.sweeps_device(w.data_ptr(), n, 8)
    if rank == 0:
        print(f"sweep pair: {s.stats()['t_sweeps']*1e3:.3f} ms", flush=True)
X = None
w = db.clone(); s.solve_device(w.data_ptr(), n, 8)
X = w.cpu().numpy().T
res = (np.linalg.norm(A @ X - B, axis=0) / np.linalg.norm(B, axis=0)).max()
err = (np.linalg.norm(X - xs, axis=0) / np.linalg.norm(xs, axis=0)).max()
print(f"rank {rank}: relres {res:.2e} err {err:.3e}", flush=True)
# distributed right-hand side (solve_ABdist): this rank's slab only
if world > 1:
    m_loc = n // world
    lo = rank * m_loc
    hi = n if rank == world - 1 else lo + m_loc
    Bl = np.asfortranarray(B[lo:hi].copy())
    for it in range(3):
        Bl[:] = B[lo:hi]
        dist.barrier()
        t0 = time.perf_counter()
        s.solve_dist(Bl, lo)
        t1 = time.perf_counter()
    errl = np.abs(Bl - X[lo:hi]).max()
    print(f"rank {rank}: solve_dist {1e3*(t1-t0):.2f} ms wall, slab equals the replicated solve: max diff {errl:.2e}", flush=True)
s.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
