"""N>1 host logic on CPU: (1) the distributed plan executed by the CPU plan interpreter with
P ranks in lockstep gives the single-rank answer; (2) world_size-2 gloo processes derive the
same partition / transfer schedule independently (what the NCCL path relies on)."""
import ctypes
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, synth_case

P = ctypes.POINTER


def _ip(a):
    return a.ctypes.data_as(P(ctypes.c_int)) if a is not None else None


def _dp(a):
    return a.ctypes.data_as(P(ctypes.c_double)) if a is not None else None


@pytest.mark.parametrize("nranks,outer", [(2, 8), (3, 8), (4, 8), (8, 8), (2, 1), (4, 1), (8, 1), (8, 2)])
def test_lockstep_ranks_match_single_rank(sim_lib, nranks, outer, monkeypatch):
    """outer = 1 / 2 makes the distribution blocks of the top fronts 64 / 128 columns wide, so that the small test
    grid exercises several owner rotations, look-ahead panels and update-matrix column blocks per top front."""
    monkeypatch.setenv("NKP_OUTER", str(outer))
    c = synth_case(24, 28, 16, seed=2)
    n = c["n"]
    A = sp.csr_matrix((c["nzval"], c["colind"], c["rowptr"]), shape=(n, n))
    A = (A - 50.0 * sp.eye(n)).tocsr()
    A.sort_indices()
    rp, ci, nz = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
    xs = np.random.default_rng(0).standard_normal((n, 2))
    B = np.asfortranarray(A @ xs)
    out = []
    for nr in (1, nranks):
        X = np.zeros_like(B, order="F")
        stats = np.zeros(8)
        part = np.zeros(4 * nr)
        rc = sim_lib.nkp_sim_run_dist(n, _ip(rp), _ip(ci), _dp(nz), _ip(c["i"]), _ip(c["j"]), _ip(c["k"]), 64, 48,
                                      _dp(B), 2, _dp(X), _dp(stats), None, 0, nr, _dp(part))
        assert rc == 0, rc
        out.append((X, stats.copy(), part.reshape(nr, 4).copy()))
    X1, st1, _ = out[0]
    Xp, stp, part = out[1]
    assert np.array_equal(X1, Xp)                      # same arithmetic, same order: bitwise equal
    assert np.linalg.norm(Xp - xs) / np.linalg.norm(xs) <= 1e-12
    assert part[:, 0].sum() == st1[0]                  # every front has exactly one owner
    assert part[:, 1].sum() == part[:, 2].sum() > 0    # sends match receives
    assert abs(part[:, 3].sum() - st1[5]) <= 1e-6 * st1[5]   # flops partition the total
    assert part[:, 3].max() <= 0.85 * st1[5]           # and no rank does (nearly) everything
    # shape of the distributed top of the tree, as rank 0 plans it
    local = np.zeros(8)
    owner = np.zeros(8192, dtype=np.int32)
    xfer = np.zeros(3 * 8192, dtype=np.int32)
    nx = ctypes.c_int()
    nf = sim_lib.nkp_sim_partition(n, _ip(rp), _ip(ci), _ip(c["i"]), _ip(c["j"]), _ip(c["k"]), 64, 48, 0, nranks,
                                   _ip(owner), 8192, _ip(xfer), 8192, ctypes.byref(nx), _dp(local))
    assert nf > 0
    assert local[4] >= 1 and local[7] == nranks                    # the root is a top front and its group is everyone
    if outer == 1:
        assert local[5] >= 3                                        # several distribution blocks per top front


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libnkp_sim.so"))
    c = synth_case(20, 24, 10, seed=1)
    owner = np.zeros(4096, dtype=np.int32)
    xfer = np.zeros(3 * 4096, dtype=np.int32)
    nx = ctypes.c_int()
    local = np.zeros(8)
    nf = lib.nkp_sim_partition(c["n"], _ip(c["rowptr"]), _ip(c["colind"]), _ip(c["i"]), _ip(c["j"]), _ip(c["k"]), 64, 48,
                               rank, world, _ip(owner), 4096, _ip(xfer), 4096, ctypes.byref(nx), _dp(local))
    mine = dict(nf=nf, owner=owner[:nf].tolist(), xfers=xfer[:3 * nx.value].reshape(-1, 3).tolist(),
                flops_local=local[0], flops=local[1])
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = all(g["owner"] == gathered[0]["owner"] and g["xfers"] == gathered[0]["xfers"] for g in gathered)
    tot = sum(g["flops_local"] for g in gathered)
    ok = ok and abs(tot - mine["flops"]) <= 1e-6 * mine["flops"] and nf > 0
    ok = ok and set(gathered[0]["owner"]) == set(range(world))
    ok = ok and all(s != d for _, s, d in gathered[0]["xfers"]) and len(gathered[0]["xfers"]) > 0
    if rank == 0:
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_partition_consistent(sim_lib):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok
    assert all(p.exitcode == 0 for p in procs)
