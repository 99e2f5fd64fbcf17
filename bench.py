#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: numeric factor time (s) & solves/sec of the
tracer-Jacobian preconditioner on synthetic POP-style grids.

A "step" is one pass of the hot path on one batch of synthetic input: ONE numeric
refactorisation (new values, same sparsity pattern, analysis reused -- a Newton step)
followed by ONE batched solve of NRHS tracers with iterative refinement.

  python bench.py --gpus N --steps K --warmup W            (own arm, CUDA path)
  python bench.py --impl reference --gpus N --steps K ...  (CPU SuperLU stand-in, rank 0 only)

`value`  : seconds per numeric factorisation, operand values already resident in HBM,
           measured with CUDA events on the solver's own stream.
`e2e`    : the same through the reference-facing C ABI with HOST buffers (nkp_factor /
           nkp_solve: H2D of the values and right-hand sides, D2H of the solutions inside
           the timed region).
Every rank checks its own answer (relative residual <= 1e-10, manufactured solution
recovered to SOL_TOL) and the run exits non-zero if any rank fails: rc == 0 at N GPUs is
parity evidence.  Accuracy keys come first in the JSON line, then the second half of the
metric (solves_per_sec), then the rooflines.

The own arm's default workload is the shape the metric is quoted on (gx1v6).  The CPU arm
cannot factor that shape in minutes (9e13 flop on one SuperLU core), so `--impl reference`
MEASURES the gx3v7 shape; the own arm carries the same gx3v7 measurement as `secondary`, and
both arms time the 64x74x38 sample, so that measured GPU/CPU pairs on common inputs exist
in every run.  Anything scaled by a flop model is labelled "extrapolated".
"""
from __future__ import annotations

import argparse
import faulthandler
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (imt, jmt, km, option set, BASELINE.json config it stands for)
    "small": (20, 24, 10, "min", "configs[0] synthetic 20x24x10"),
    "sample": (64, 74, 38, "min", "bounded CPU sample 64x74x38 of the same generator"),
    "gx3v7": (100, 116, 60, "min", "configs[1] gx3v7-shape 100x116x60"),
    "gx1v6": (320, 384, 60, "min", "configs[3] gx1v6-shape 320x384x60 (the shape the metric is quoted on)"),
    # the reference's own option set (test/test_gen_A.csh:21-24): operand written by the unchanged gen_A
    "small_ref": (20, 24, 10, "ref", "configs[0] shape, reference test options"),
    "gx3v7_ref": (100, 116, 60, "ref", "configs[1] gx3v7-shape, reference test options (upwind3 + isop_file + vmix file)"),
    "gx1v6_ref": (320, 384, 60, "ref", "configs[3] gx1v6-shape, reference test options; needs 8 GPUs (257 GB of factors)"),
}
OPTION_TEXT = {
    "min": "adv centered, hmix const, vmix const, sink const_shallow 365 10e2 (7-point rows; numpy restatement of gen_A, bit-exact)",
    "ref": "adv upwind3, hmix isop_file, vmix file, sink const_shallow 365 10e2, day_cnt 365 (test/test_gen_A.csh:21-24; "
           "<= 21-point rows; matrix file written by the reference's unchanged gen_A in the bench set-up)",
}
DEFAULT_WORKLOAD = os.environ.get("NKP_BENCH_WORKLOAD", "gx1v6")
REFERENCE_WORKLOAD = "gx3v7"   # what the CPU arm measures (see module docstring)
NRHS = 8  # BASELINE.json configs[2]: 8 tracers as batched right-hand sides
RES_TOL = 1e-10   # BASELINE.json: ||Ax-b|| / ||b||
SOL_TOL = 1e-8    # BASELINE.json: solution relative difference
# gx1v6-shape operands have cond(A) ~ 1e8 (one-year step of a transport operator damped only by a surface sink).  The
# manufactured right-hand side is rounded to double (half an ulp per entry even when formed in extended precision),
# and cond(A) turns that into 0.85e-8 .. 3.9e-8 (depending on x*) of distance between x* and the EXACT solution of the rounded system --
# for any solver.  The extra-precise refinement converges to that exact solution (the iterates stop changing at the
# 1e-16 level, profiles/r02_refine_probe_gx1v6.log; two different factorisations agree to 1e-10,
# tests/test_gpu_parity.py::test_full_size_gx1v6_properties), so what is asserted against x* at this shape is the floor.
SOL_TOL_GX1V6 = 1e-7


def build_case(name, seed=1):
    """Operand of a workload.  "min" option set: synth.assemble_crs (numpy restatement of gen_sparse_matrix,
    bit-exact against gen_A in tests/).  "ref" option set: the circulation file is synthesised, the matrix
    file is written by the reference's own unchanged generator (oracle/_ref/gen_A, the input generator --
    not the checker) and read back."""
    from nk_ocn_tracer_jacobian_precond_b200 import synth
    imt, jmt, km, optset = WORKLOADS[name][:4]
    g = synth.make_grid(imt, jmt, km, seed=seed)
    c = synth.make_circulation(g, seed=seed)
    if optset == "min":
        n, rp, ci, nz, (ii, jj, kk, _) = synth.assemble_crs(g, c)
    else:
        # one generation per node: local rank 0 writes the matrix file, the other ranks of a multi-GPU run read it
        cache = os.path.join(tempfile.gettempdir(), f"nkp_case_{name}_seed{seed}.nc")
        if int(os.environ.get("LOCAL_RANK", "0")) == 0 and not os.path.exists(cache):
            gen_a = os.path.join(ROOT, "oracle", "_ref", "gen_A")
            if not os.path.exists(gen_a):
                raise RuntimeError("oracle/_ref/gen_A is missing (built by __graft_entry__.build() where /root/reference exists)")
            full = synth.make_full_fields(g, c, seed=seed)
            with tempfile.TemporaryDirectory(prefix="nkp_bench_") as td:
                circ = os.path.join(td, "circ.nc")
                synth.write_circ_file(circ, g, c, full)
                del full
                open(os.path.join(td, "opts.txt"), "w").write(synth.REFTEST_OPTS.format(circ=circ))
                subprocess.check_call([gen_a, "-o", os.path.join(td, "opts.txt"), os.path.join(td, "A.nc")],
                                      stdout=subprocess.DEVNULL)
                os.replace(os.path.join(td, "A.nc"), cache + ".part")
            os.replace(cache + ".part", cache)
        t_wait = time.time()
        while not os.path.exists(cache):
            if time.time() - t_wait > 1800:
                raise RuntimeError(f"{cache} did not appear (local rank 0 generates it)")
            time.sleep(1.0)
        m = synth.read_matrix_file(cache)
        rp, ci, nz = m["rowptr"].astype(np.int32), m["colind"].astype(np.int32), m["nzval_row_wise"].astype(np.float64)
        ii, jj, kk = (m["tracer_state_ind_to_" + q].astype(np.int32) for q in "ijk")
        n = len(rp) - 1
    return dict(n=n, rowptr=rp, colind=ci, nzval=nz, coords=(ii, jj, kk), shape=(imt, jmt, km), name=name,
                desc=WORKLOADS[name][4], options=OPTION_TEXT[optset])


def spmv_extended(rowptr, colind, nzval, X):
    """B = A X with products and sums in x87 extended precision (64-bit significand), rounded once to
    double: the manufactured right-hand side then carries half an ulp of error per entry instead of the
    (row length) ulps of a working-precision SpMV -- the solution error against X measures the solver,
    not the way b was formed."""
    import scipy.sparse as sp
    n = len(rowptr) - 1
    Al = sp.csr_matrix((nzval.astype(np.longdouble), colind, rowptr), shape=(n, n))
    return np.asfortranarray((Al @ X.astype(np.longdouble)).astype(np.float64))


def measured_peaks():
    peaks = {}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks.update(json.load(open(p)))
        peaks["_hbm_src"] = "MEASURED_PEAKS.json"
    else:
        peaks["hbm_gbs"] = 6650.0
        peaks["_hbm_src"] = "fallback (B200_PROFILING.md)"
    # FP64: MEASURED_PEAKS.json has no FP64 entry; use this repo's own cuBLAS DGEMM measurement
    q = os.path.join(ROOT, "profiles", "r01_dgemm_peak.json")
    if os.path.exists(q):
        d = json.load(open(q))
        peaks["fp64_tflops"] = d["fp64_dgemm_tflops_sustained"]
        peaks["_fp64_src"] = "profiles/r01_dgemm_peak.json (cuBLAS DGEMM 8192^3 on this pool's B200)"
    else:
        peaks["fp64_tflops"] = 37.0
        peaks["_fp64_src"] = "nominal B200 FP64"
    return peaks


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, device=0):
        super().__init__(daemon=True)
        self.device = device
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port (scipy's serial SuperLU), used ONLY as the reported baseline
# ------------------------------------------------------------------------------------------------

def _sim_analysis(case):
    """Nested-dissection permutation and algorithmic flop count of the GPU path's analysis (CPU only, through the
    plan interpreter's build of the analysis code)."""
    import ctypes
    sim = os.path.join(ROOT, "oracle", "libnkp_sim.so")
    if not os.path.exists(sim):
        return None, None
    lib = ctypes.CDLL(sim)
    P = ctypes.POINTER
    n = case["n"]
    stats = np.zeros(8)
    perm = np.zeros(n, dtype=np.int32)
    arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (case["rowptr"], case["colind"], *case["coords"])]
    ip = lambda a: a.ctypes.data_as(P(ctypes.c_int))
    rc = lib.nkp_sim_run(n, ip(arrs[0]), ip(arrs[1]), case["nzval"].ctypes.data_as(P(ctypes.c_double)), ip(arrs[2]),
                         ip(arrs[3]), ip(arrs[4]), 64, 96, None, 0, None, stats.ctypes.data_as(P(ctypes.c_double)),
                         ip(perm), 1)
    return (perm, float(stats[5])) if rc == 0 else (None, None)


def cpu_factor_solve(case, variant, nrhs=NRHS):
    """One CPU factorisation + nrhs solves of `case` with scipy's serial SuperLU.
    variant "nd_static": the GPU path's nested-dissection order applied symmetrically, NATURAL column order,
    diag_pivot_thresh = 0 -- like-for-like ordering and static pivoting, the
    closest serial analogue of the reference's ParMETIS + static-pivoting SuperLU_DIST run;
    variant "colamd": library defaults (COLAMD, partial pivoting; (i))."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    n = case["n"]
    A = sp.csr_matrix((case["nzval"], case["colind"], case["rowptr"]), shape=(n, n))
    flops = None
    if variant == "nd_static":
        perm, flops = _sim_analysis(case)
        if perm is None:
            raise RuntimeError("oracle/libnkp_sim.so missing")
        iperm = np.empty(n, dtype=np.int64)
        iperm[perm] = np.arange(n)
        M = A[iperm][:, iperm].tocsc()
        t0 = time.perf_counter()
        lu = spla.splu(M, permc_spec="NATURAL", diag_pivot_thresh=0.0, options=dict(SymmetricMode=False))
        t_factor = time.perf_counter() - t0
    else:
        iperm = None
        M = A.tocsc()
        t0 = time.perf_counter()
        lu = spla.splu(M, permc_spec="COLAMD")
        t_factor = time.perf_counter() - t0
    rng = np.random.default_rng(0)
    xs = rng.standard_normal((n, nrhs))
    B = A @ xs
    Bp = B[iperm] if iperm is not None else B
    t0 = time.perf_counter()
    Xp = np.column_stack([lu.solve(Bp[:, c]) for c in range(nrhs)])
    t_solve = time.perf_counter() - t0
    X = np.empty_like(Xp)
    if iperm is not None:
        X[iperm] = Xp
    else:
        X = Xp
    relres = float((np.linalg.norm(A @ X - B, axis=0) / np.linalg.norm(B, axis=0)).max())
    return {"variant": variant, "n": n, "factor_s": t_factor, "solve_s_per_rhs": t_solve / nrhs,
            "solves_per_sec": nrhs / t_solve, "nnz_lu": int(lu.L.nnz + lu.U.nnz), "relres_no_refinement": relres,
            "factor_flops_model": flops, "gflops": flops / t_factor * 1e-9 if flops else None}


def cpu_baseline(full_flops):
    """cpu_baseline of the own arm: the oracle port on the bounded 64x74x38 sample (about 10 s of one core).
    `value` is the MEASURED factor time of the sample; the figure scaled to the workload by the factor-flop ratio
    is kept only under "extrapolated"."""
    c = build_case("sample")
    nd = cpu_factor_solve(c, "nd_static")
    co = cpu_factor_solve(c, "colamd", nrhs=2)
    scale = full_flops / nd["factor_flops_model"] if nd["factor_flops_model"] else None
    return {
        "value": nd["factor_s"], "unit": "s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
        "sample": f"scipy.sparse.linalg.splu (serial SuperLU) on the 64x74x38 grid of the same generator, n={c['n']}: "
                  f"the GPU path's nested-dissection order + no pivoting: factor {nd['factor_s']:.2f} s "
                  f"({nd['gflops']:.1f} GFLOP/s), {nd['solve_s_per_rhs'] * 1e3:.1f} ms per solve; "
                  f"COLAMD + partial pivoting: factor {co['factor_s']:.2f} s.  MEASURED, not scaled; the same sample is "
                  f"timed on the GPU in `sample_on_gpu`",
        "nd_static": nd, "colamd": co,
        "extrapolated": {"factor_s_at_workload": nd["factor_s"] * scale if scale else None, "flop_ratio": scale,
                         "note": "sample time x ratio of algorithmic factor flops; NOT a measurement"},
    }


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: scipy's serial SuperLU), rank 0 only.
    Measures the gx3v7 shape for real (about 1.3e12 flop, 1-2 minutes on one core, 4 GB); one factorisation
    is one step, and as many steps are run as fit in REF_BUDGET_S seconds (`steps` reports the count)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = float(os.environ.get("NKP_REF_BUDGET_S", "200"))
    wl = args.workload if args.workload_given else REFERENCE_WORKLOAD
    case = build_case(wl)
    t_start = time.perf_counter()
    runs = []
    want = max(1, args.steps)
    while len(runs) < want:
        runs.append(cpu_factor_solve(case, "nd_static"))
        spent = time.perf_counter() - t_start
        if spent + 1.2 * spent / len(runs) > budget:
            break
    v = float(np.mean([r["factor_s"] for r in runs]))
    sps = float(np.mean([r["solves_per_sec"] for r in runs]))
    colamd = None
    if time.perf_counter() - t_start + 2.0 * v < budget:
        colamd = cpu_factor_solve(case, "colamd", nrhs=2)
    line = {
        "impl": "reference", "metric": "numeric_factor_time_s", "value": v, "unit": "s", "n_gpus": args.gpus,
        "steps": len(runs), "warmup": 0, "requested_steps": args.steps, "requested_warmup": args.warmup,
        "ms_per_step": (v + NRHS / sps) * 1e3, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "solves_per_sec": sps,
        "config": {"workload": f"{case['desc']}: one numeric factorisation + {NRHS} solves per step (MEASURED at this "
                               f"shape; the own arm's `secondary` record is the same workload)",
                   "n": case["n"], "nnz": int(len(case["nzval"])), "nrhs": NRHS, "options": case["options"]},
        "cpu_baseline": {"value": v, "unit": "s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
                         "sample": f"the whole {wl} workload, {len(runs)} factorisation(s): scipy.sparse.linalg.splu (serial "
                                   f"SuperLU -- the reference pins SuperLU_DIST 5.1.3, unbuildable here) with the GPU path's "
                                   f"nested-dissection order and static pivoting, 1 core",
                         "runs": runs, "colamd_partial_pivoting": colamd},
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "solves_per_sec": sps},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------

class Ctx:
    pass


def measure(ctx, case, steps, warmup, detail):
    """Device-resident and end-to-end timings of one workload on the ranks of ctx; every rank checks its answer."""
    import scipy.sparse as sp
    import torch
    from nk_ocn_tracer_jacobian_precond_b200 import solver
    dist, dev, rank, world = ctx.dist, ctx.dev, ctx.rank, ctx.world
    n, nnz = case["n"], len(case["nzval"])
    comm = None
    if world > 1:
        # subtree-to-GPU sharding: NCCL communicator of the solver, id shipped through torch.distributed
        uid = [solver.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = (rank, world, uid[0])
    s = solver.TracerJacobianSolver(n, case["rowptr"], case["colind"], coords=case["coords"], comm=comm, device=ctx.local)
    st0 = s.stats()

    nsteps = warmup + steps
    rng = np.random.default_rng(1234)   # identical operands on every rank
    # Newton sequence: new values every step, same pattern (BASELINE.json configs[4])
    host_vals = [case["nzval"] * (1.0 + 1e-3 * rng.standard_normal(nnz)) for _ in range(min(nsteps, 3))]
    xs = rng.standard_normal((n, NRHS))
    host_B = [spmv_extended(case["rowptr"], case["colind"], v, xs) for v in host_vals]
    dev_vals = [torch.tensor(v, device=dev) for v in host_vals]
    dev_B = [torch.tensor(np.ascontiguousarray(b.T), device=dev) for b in host_B]   # (nrhs, n) row-major == column-major n x nrhs
    work_B = torch.empty_like(dev_B[0])

    def barrier():
        s.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    sol_tol = ctx.sol_tol if ctx.sol_tol is not None else (SOL_TOL_GX1V6 if case["shape"][0] >= 320 else SOL_TOL)

    def check(X, k, what, strict=True):
        A = sp.csr_matrix((host_vals[k], case["colind"], case["rowptr"]), shape=(n, n))
        relres = float((np.linalg.norm(A @ X - host_B[k], axis=0) / np.linalg.norm(host_B[k], axis=0)).max())
        solerr = float((np.linalg.norm(X - xs, axis=0) / np.linalg.norm(xs, axis=0)).max())
        ok = np.isfinite(relres) and relres <= RES_TOL and (solerr <= sol_tol or not strict)
        if not ok:
            print(f"[bench] rank {rank}: PARITY FAILURE ({what}, {case['name']}): relres {relres:.3e} (tol {RES_TOL:g}), "
                  f"solution error {solerr:.3e} (tol {sol_tol:g})", file=sys.stderr, flush=True)
            ctx.failed = True
        return relres, solerr

    # ---------------- device-resident arm -------------------------------------------------
    s.set_profile(detail)
    fact_t, solve_t, gemm_t, refine = [], [], [], []
    launches0 = None
    sampler = None
    berr = np.zeros(NRHS)
    for it in range(nsteps):
        k = it % len(dev_vals)
        if it == warmup:
            barrier()
            launches0 = s.stats()["kernel_launches"]
            if rank == 0 and detail:
                sampler = ClockSampler(ctx.local)
                sampler.start()
            t_wall0 = time.perf_counter()
        s.factor_device(dev_vals[k].data_ptr())
        work_B.copy_(dev_B[k])
        torch.cuda.current_stream().synchronize()
        if dist is not None:
            dist.barrier()   # ranks leave the factorisation at different times; keep that skew out of the solve timing
        berr = s.solve_device(work_B.data_ptr(), n, NRHS)
        if it >= warmup:
            st = s.stats()
            fact_t.append(st["t_factor"])
            solve_t.append(st["t_solve"])
            gemm_t.append((st["t_gemm"], st["n_gemm"], st["t_trsm"], st["t_diag"], st["t_extend_add"], st["t_scatter"]))
            refine.append(st["refine_steps"])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    st = s.stats()
    launches = st["kernel_launches"] - launches0
    k_last = (nsteps - 1) % len(dev_vals)
    relres, solerr = check(work_B.cpu().numpy().T, k_last, "device arm")

    out = {"sol_tol": sol_tol, "relres_max": relres, "solution_err_max": solerr, "berr_max": float(np.max(berr)),
           "refine_steps": float(np.mean(refine)), "tiny_pivots_replaced": int(st["tiny_pivots"])}

    def mx(v):
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    extra = {}
    if detail:
        # raw sweep pair (no refinement) for the HBM roofline of the solve
        s.set_profile(False)
        sweep_t = []
        for it in range(3 + 3):
            work_B.copy_(dev_B[0])
            torch.cuda.current_stream().synchronize()
            s.sweeps_device(work_B.data_ptr(), n, NRHS)
            if it >= 3:
                sweep_t.append(s.stats()["t_sweeps"])
        # the same batched solve with the normwise stopping rule (nkp_options.refine_rule = 1): the default
        # above is SuperLU's componentwise-berr rule, which keeps refining long after ||r|| / ||b|| is at
        # rounding level
        s.set_refine_rule(1)
        nw_t, nw_steps = [], []
        for it in range(2 + 3):
            work_B.copy_(dev_B[k_last])
            torch.cuda.current_stream().synchronize()
            if dist is not None:
                dist.barrier()
            s.solve_device(work_B.data_ptr(), n, NRHS)
            if it >= 2:
                nw_t.append(s.stats()["t_solve"])
                nw_steps.append(s.stats()["refine_steps"])
        # (this rule bounds the residual, not the error: only the residual is asserted)
        nw_relres, nw_solerr = check(work_B.cpu().numpy().T, k_last, "normwise rule", strict=False)
        s.set_refine_rule(0)
        # residual SpMV r = b - A x (refinement) and the CRS -> front scatter: device time with events
        dx = torch.randn(NRHS, n, dtype=torch.float64, device=dev)
        dr = torch.empty_like(dx)
        spmv_t = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(3 + 5):
            s.sync()
            t0 = time.perf_counter()
            s.residual_device(dx.data_ptr(), dev_B[0].data_ptr(), dr.data_ptr(), NRHS)
            s.sync()
            if it >= 3:
                spmv_t.append(time.perf_counter() - t0)
        extra = dict(sweep_s=mx(np.mean(sweep_t)), nw_solve_s=mx(np.mean(nw_t)), nw_steps=float(np.mean(nw_steps)),
                     nw_relres=nw_relres, nw_solerr=nw_solerr, spmv_s=float(np.min(spmv_t)))

    # ---------------- end-to-end arm (host buffers through the C ABI) ------------------------
    # Inputs live in page-locked host memory (the contract's "pinned host memory"); nkp_factor / nkp_solve
    # detect that and DMA straight from / to the caller's arrays.  A pageable variant is timed next to it.
    def pinned_like(a):
        t = torch.empty(a.shape[::-1] if a.ndim == 2 else a.shape, dtype=torch.float64).pin_memory()
        v = t.numpy().T if a.ndim == 2 else t.numpy()
        return t, v

    # With several GPUs the right-hand side is ROW-DISTRIBUTED like solve_ABdist's (src/solve_ABdist.c:141-144): every
    # rank passes and gets back only its slab of n / P rows (nkp_solve_dist); the slabs travel between GPUs over NVLink.
    m_loc = n // world
    lo = rank * m_loc
    hi = n if rank == world - 1 else lo + m_loc
    e2e = {}
    for mode in (("pinned", "pageable") if detail else ("pinned",)):
        slab = np.empty((hi - lo, NRHS), order="F")
        if mode == "pinned":
            keep_v, vals_h = zip(*[pinned_like(v) for v in host_vals])
            for dst, src in zip(vals_h, host_vals):
                dst[:] = src
            keep_b, Bh = pinned_like(slab)
        else:
            vals_h, Bh = host_vals, slab
        e2e_f, e2e_s = [], []
        wu = min(warmup, 2)
        e2e_berr = np.zeros(NRHS)
        for it in range(wu + steps):
            k = it % len(host_vals)
            Bh[:] = host_B[k][lo:hi]
            barrier()
            t0 = time.perf_counter()
            s.factor(vals_h[k])
            s.sync()
            if dist is not None:
                dist.barrier()
            t1 = time.perf_counter()
            e2e_berr = s.solve_dist(Bh, lo)
            t2 = time.perf_counter()
            if it >= wu:
                e2e_f.append(t1 - t0)
                e2e_s.append(t2 - t1)
        if world == 1:
            check(np.array(Bh), (wu + steps - 1) % len(host_vals), f"e2e arm ({mode} host buffers)")
        else:
            # the distributed solution against the manufactured one (global 2-norm per column, summed over the slabs of all
            # ranks) and the backward error of the whole system
            sq = torch.tensor(np.stack([((Bh - xs[lo:hi]) ** 2).sum(axis=0), (xs[lo:hi] ** 2).sum(axis=0)]), device=dev)
            dist.all_reduce(sq, op=dist.ReduceOp.SUM)
            dist_err = float(torch.sqrt(sq[0] / sq[1]).max().item())
            if not (dist_err <= sol_tol and e2e_berr.max() <= 1e-14):
                print(f"[bench] rank {rank}: PARITY FAILURE (e2e arm, {mode}, row-distributed): solution error {dist_err:.3e} "
                      f"(tol {sol_tol:g}), berr {e2e_berr.max():.3e}", file=sys.stderr, flush=True)
                ctx.failed = True
        e2e[mode] = (mx(np.mean(e2e_f)), mx(np.mean(e2e_s)))

    factor_s = mx(np.mean(fact_t))
    solve_s = mx(np.mean(solve_t))
    out.update({
        "factor_s": factor_s, "solve_s": solve_s, "solves_per_sec": NRHS / solve_s,
        "factor_tflops": st["factor_flops"] / factor_s * 1e-12,
        "step_ms": mx(t_wall / steps * 1e3),
        "e2e_factor_s": e2e["pinned"][0], "e2e_solve_s": e2e["pinned"][1], "e2e_solves_per_sec": NRHS / e2e["pinned"][1],
        "launches": int(launches), "clocks": clocks, "n": n, "nnz": nnz, "stats": st, "analysis_s": st0["t_analysis"],
        "gemm_t": np.array(gemm_t), "fact_mean_local": float(np.mean(fact_t)),
    })
    if "pageable" in e2e:
        out["e2e_pageable_factor_s"], out["e2e_pageable_solve_s"] = e2e["pageable"]
    out.update(extra)
    s.close()
    return out


def compact(m, case):
    """Short record of a secondary workload."""
    return {"workload": case["desc"], "n": m["n"], "nnz": m["nnz"], "nrhs": NRHS, "options": case["options"],
            "relres_max": m["relres_max"], "solution_err_max": m["solution_err_max"], "berr_max": m["berr_max"],
            "refine_steps": m["refine_steps"], "value": m["factor_s"], "unit": "s", "factor_tflops": m["factor_tflops"],
            "solve_s": m["solve_s"], "solves_per_sec": m["solves_per_sec"],
            "e2e": {"value": m["e2e_factor_s"], "unit": "s", "solve_s": m["e2e_solve_s"],
                    "solves_per_sec": m["e2e_solves_per_sec"]},
            "factor_flops": m["stats"]["factor_flops"], "nnz_lu": m["stats"]["nnz_lu"], "analysis_s": m["analysis_s"]}


def run_own(args):
    import torch

    ctx = Ctx()
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    ctx.dist = None
    ctx.failed = False
    ctx.sol_tol = args.sol_tol
    if ctx.world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(ctx.local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local))
        ctx.dist = dist
    ctx.dev = torch.device("cuda", ctx.local)
    torch.cuda.set_device(ctx.dev)
    rank, world = ctx.rank, ctx.world

    wl = args.workload
    t0 = time.perf_counter()
    case = build_case(wl)
    t_gen = time.perf_counter() - t0
    m = measure(ctx, case, args.steps, args.warmup, detail=True)
    st, n, nnz = m["stats"], m["n"], m["nnz"]

    secondary = None
    sample_gpu = None
    if not args.no_secondary:
        if wl != REFERENCE_WORKLOAD:
            c2 = build_case(REFERENCE_WORKLOAD)
            secondary = compact(measure(ctx, c2, 3, 2, detail=False), c2)
        c3 = build_case("sample")
        sample_gpu = compact(measure(ctx, c3, 3, 2, detail=False), c3)

    # every rank reports; any failure anywhere fails the run
    fail = torch.tensor([1.0 if ctx.failed else 0.0], device=ctx.dev)
    if ctx.dist is not None:
        ctx.dist.all_reduce(fail, op=ctx.dist.ReduceOp.MAX)
    any_failed = bool(fail.item() > 0)

    if rank == 0:
        peaks = measured_peaks()
        g = m["gemm_t"]
        t_gemm, n_gemm = float(g[:, 0].mean()), float(g[:, 1].mean())
        gemm_tflops = st["gemm_flops"] / t_gemm * 1e-12 if t_gemm > 0 else None
        roofline = {
            "kernel": "k_gemm (Schur-complement update, FP64 DMMA)", "bound": "tensor",
            "achieved": gemm_tflops, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s",
            "frac": gemm_tflops / peaks["fp64_tflops"] if gemm_tflops else None,
            "peak_source": peaks["_fp64_src"], "traffic": None,
            "launches_per_factor": n_gemm, "avg_launch_ms": t_gemm / n_gemm * 1e3 if n_gemm else None,
            "alg_flops_per_launch": st["gemm_flops"] / n_gemm if n_gemm else None,
            "share_of_factor_time": t_gemm / m["fact_mean_local"],
            "factor_breakdown_s": {"gemm": t_gemm, "trsm": float(g[:, 2].mean()), "diag": float(g[:, 3].mean()),
                                   "extend_add": float(g[:, 4].mean()), "zero+scatter": float(g[:, 5].mean())},
            "factor_overall_tflops": m["factor_tflops"],
        }
        solve_gbs = st["solve_bytes"] / m["sweep_s"] * 1e-9
        roofline_solve = {
            "kernel": "k_sweep_big + k_fwd_small/k_bwd_small sweep pair (nrhs=%d, no refinement)" % NRHS, "bound": "hbm",
            "achieved": solve_gbs, "peak": peaks["hbm_gbs"] * world, "unit": "GB/s",
            "frac": solve_gbs / (peaks["hbm_gbs"] * world),
            "peak_source": peaks["_hbm_src"] + (f" x {world} GPUs (whole-job bytes over the max-over-ranks time)" if world > 1 else ""),
            "traffic": None, "sweep_pair_ms": m["sweep_s"] * 1e3,
            "alg_bytes_per_sweep_pair": st["solve_bytes"],
        }
        spmv_bytes = 12.0 * nnz + 4.0 * (n + 1) + 24.0 * n * NRHS   # ONE pass over A for all rhs
        roofline_spmv = {
            "kernel": "k_residual (r = b - A x, %d rhs in one pass over A; host-timed launch+sync)" % NRHS, "bound": "hbm",
            "achieved": spmv_bytes / m["spmv_s"] * 1e-9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": spmv_bytes / m["spmv_s"] * 1e-9 / peaks["hbm_gbs"], "ms": m["spmv_s"] * 1e3, "traffic": None,
            "alg_bytes": spmv_bytes,
        }
        scatter_s = float(g[:, 5].mean())
        roofline_scatter = {
            "kernel": "equilibration + zero-fill of the factor arena + k_scatter (CRS -> fronts)", "bound": "hbm",
            "alg_bytes": 20.0 * nnz, "ms": scatter_s * 1e3,
            "achieved": 20.0 * nnz / scatter_s * 1e-9 if scatter_s > 0 else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": 20.0 * nnz / scatter_s * 1e-9 / peaks["hbm_gbs"] if scatter_s > 0 else None,
            "note": "the timed region also clears the whole factor arena (8 B per factor entry), which dominates it",
        }
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tpath) and world == 1:
            tj = json.load(open(tpath))
            if wl in tj:
                roofline["traffic"] = tj[wl].get("k_gemm_bytes_per_launch")
                roofline["traffic_source"] = tj["source"]
                roofline_solve["traffic"] = tj[wl].get("sweep_pair_bytes")
                roofline_spmv["traffic"] = tj[wl].get("k_residual_bytes_per_launch")
        cb = cpu_baseline(st["factor_flops"]) if not args.no_cpu else None
        if cb and sample_gpu:
            cb["sample_on_gpu"] = {"factor_s": sample_gpu["value"], "e2e_factor_s": sample_gpu["e2e"]["value"],
                                   "solves_per_sec": sample_gpu["solves_per_sec"],
                                   "measured_cpu_over_gpu_factor": cb["value"] / sample_gpu["e2e"]["value"]}
        line = {
            "metric": "numeric_factor_time_s", "value": m["factor_s"], "unit": "s",
            "relres_max": m["relres_max"], "solution_err_max": m["solution_err_max"], "berr_max": m["berr_max"],
            "refine_steps": m["refine_steps"], "solves_per_sec": NRHS / m["solve_s"], "solve_s": m["solve_s"],
            "parity_checked_on_every_rank": True, "parity_tolerances": {"relres": RES_TOL, "solution_err": m["sol_tol"],
                                  "note": "solution_err is against the manufactured x*; at gx1v6-shape cond(A) ~ 1e8 and the "
                                          "rounding of b alone moves the exact solution by ~1e-8 (see SOL_TOL_GX1V6 in bench.py)"},
            "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["step_ms"], "higher_is_better": False,
            "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{case['desc']}: one numeric refactorisation + one batched solve of {NRHS} tracers per step",
                "n": n, "nnz": nnz, "nrhs": NRHS, "options": case["options"],
                "rhs": "b = A x* with x* ~ N(0,1), formed in x87 extended precision and rounded once",
                "nnz_lu": st["nnz_lu"], "factor_flops": st["factor_flops"], "fronts": st["n_fronts"],
                "levels": st["n_levels"], "max_front": st["max_front"], "analysis_s": m["analysis_s"],
                "l2": "operand and factors exceed L2 (%.1f GB heap); no flush needed" % (st["heap_bytes"] * 1e-9),
                "parallelism": "1 GPU" if world == 1 else
                f"{world} GPUs: nested-dissection subtrees sharded over the GPUs; the top separator fronts are factored by the "
                f"group of GPUs below them (block-cyclic column blocks, panel broadcasts over NCCL with look-ahead) and swept "
                f"by every member, so a sweep pair needs {int(st['n_xfers'])} update-vector transfers and one solution exchange",
                "generate_s": t_gen,
            },
            "tiny_pivots_replaced": m["tiny_pivots_replaced"],
            "solve_normwise_rule": {"solve_s": m["nw_solve_s"], "solves_per_sec": NRHS / m["nw_solve_s"],
                                    "refine_steps": m["nw_steps"], "relres_max": m["nw_relres"],
                                    "solution_err_max": m["nw_solerr"],
                                    "rule": "stop when ||b - A x||_2 <= 1e-14 ||b||_2 (nkp_options.refine_rule = 1)"},
            "roofline": roofline, "roofline_solve": roofline_solve, "roofline_spmv": roofline_spmv,
            "roofline_scatter": roofline_scatter, "cpu_baseline": cb,
            "e2e": {"value": m["e2e_factor_s"], "unit": "s", "h2d_bytes_per_step": 8 * nnz + 8 * n * NRHS,
                    "d2h_bytes_per_step": 8 * n * NRHS, "bytes_note": "whole job; with N GPUs every rank moves its n / N rows of B and X "
                    "(nkp_solve_dist) and all ranks upload the values", "solve_s": m["e2e_solve_s"], "solves_per_sec": NRHS / m["e2e_solve_s"],
                    "host_buffers": "page-locked (DMA straight from / to the caller's arrays)",
                    "pageable_host_buffers": {"value": m.get("e2e_pageable_factor_s"), "solve_s": m.get("e2e_pageable_solve_s"),
                                              "note": "staged through the solver's pinned area with a multi-threaded copy"}},
            "secondary": secondary, "sample_on_gpu": sample_gpu,
            "gpu_launches": m["launches"], "clocks": m["clocks"],
        }
        if any_failed:
            line["parity_failed"] = True
        print(json.dumps(line))
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()
    if any_failed:
        sys.exit(3)


def main():
    faulthandler.enable()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--sol-tol", type=float, default=None, help="override the per-workload solution-error tolerance")
    ap.add_argument("--no-secondary", action="store_true", help="skip the gx3v7 / sample records of the own arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.workload_given = args.workload is not None
    if args.workload is None:
        args.workload = DEFAULT_WORKLOAD
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
