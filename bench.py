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
           nkp_solve: pinned staging + H2D of the values and right-hand sides, D2H of the
           solutions inside the timed region).
Extra keys carry the second half of the metric (solves_per_sec) and the rooflines.
"""
from __future__ import annotations

import argparse
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (imt, jmt, km, BASELINE.json config it stands for)
    "small": (20, 24, 10, "configs[0] synthetic 20x24x10"),
    "gx3v7": (100, 116, 60, "configs[1] gx3v7-shape 100x116x60"),
    "gx1v6": (320, 384, 60, "configs[3] gx1v6-shape 320x384x60 (the shape the metric is quoted on)"),
}
DEFAULT_WORKLOAD = os.environ.get("NKP_BENCH_WORKLOAD", "gx1v6")
NRHS = 8  # BASELINE.json configs[2]: 8 tracers as batched right-hand sides
CPU_SAMPLE = (64, 74, 38)  # bounded CPU sample (sub-grid of the same generator), ~10-30 s of serial SuperLU


def build_case(name, seed=1):
    from nk_ocn_tracer_jacobian_precond_b200 import synth
    imt, jmt, km = WORKLOADS[name][:3] if name in WORKLOADS else name
    g = synth.make_grid(imt, jmt, km, seed=seed)
    c = synth.make_circulation(g, seed=seed)
    n, rp, ci, nz, (ii, jj, kk, _) = synth.assemble_crs(g, c)
    return dict(n=n, rowptr=rp, colind=ci, nzval=nz, coords=(ii, jj, kk), shape=(imt, jmt, km))


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


def cpu_baseline(full_flops, nrhs=NRHS, shape=CPU_SAMPLE):
    """CPU stand-in of the reference's solver (oracle port: scipy's serial SuperLU, the
    library family the reference pins) on a bounded sample, scaled to the workload by the
    ratio of algorithmic factor flops (same nested-dissection flop model for both sizes)."""
    import ctypes
    from oracle import oracle_solve
    c = build_case(shape)
    n = c["n"]
    t0 = time.perf_counter()
    lu = oracle_solve.factor(n, c["rowptr"], c["colind"], c["nzval"])
    t_factor = time.perf_counter() - t0
    rng = np.random.default_rng(0)
    B = rng.standard_normal((n, nrhs))
    t0 = time.perf_counter()
    for col in range(nrhs):
        lu.solve(B[:, col])
    t_solve = (time.perf_counter() - t0) / nrhs
    # flop model of the sample from the plan interpreter's analysis (oracle/, CPU only)
    sample_flops = None
    sim = os.path.join(ROOT, "oracle", "libnkp_sim.so")
    if os.path.exists(sim):
        lib = ctypes.CDLL(sim)
        P = ctypes.POINTER
        stats = np.zeros(8)
        ip = lambda a: np.ascontiguousarray(a, dtype=np.int32).ctypes.data_as(P(ctypes.c_int))
        ii, jj, kk = (np.ascontiguousarray(a, dtype=np.int32) for a in c["coords"])
        rp = np.ascontiguousarray(c["rowptr"], dtype=np.int32)
        ci = np.ascontiguousarray(c["colind"], dtype=np.int32)
        os.environ.pop("NKP_SIM_REFINE", None)
        perm = np.zeros(n, dtype=np.int32)
        rc = lib.nkp_sim_run(n, ip(rp), ip(ci), c["nzval"].ctypes.data_as(P(ctypes.c_double)), ip(ii), ip(jj), ip(kk),
                             64, 96, None, 0, None, stats.ctypes.data_as(P(ctypes.c_double)), ip(perm), 1)
        if rc == 0:
            sample_flops = float(stats[5])
    # like-for-like variant: the GPU path's nested-dissection ordering applied first,
    # then SuperLU with NATURAL column order and no pivoting (diag_pivot_thresh = 0) -- same fill, same
    # flops as the GPU factorisation, on one host core
    nd = {}
    if sample_flops:
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        A = sp.csr_matrix((c["nzval"], c["colind"], c["rowptr"]), shape=(n, n))
        iperm = np.empty(n, dtype=np.int64)
        iperm[perm] = np.arange(n)
        Ap = A[iperm][:, iperm].tocsc()
        t0 = time.perf_counter()
        lu2 = spla.splu(Ap, permc_spec="NATURAL", diag_pivot_thresh=0.0, options=dict(SymmetricMode=False))
        t_nd = time.perf_counter() - t0
        nd = {"sample_factor_s_nd_order_no_pivoting": t_nd, "sample_nnz_lu_nd": int(lu2.L.nnz + lu2.U.nnz),
              "sample_gflops_nd": sample_flops / t_nd * 1e-9}
    scale = (full_flops / sample_flops) if (sample_flops and full_flops) else None
    return {
        **nd,
        "value": t_factor * scale if scale else None, "unit": "s", "cores": 1, "kind": "port",
        "sample": f"scipy.sparse.linalg.splu (serial SuperLU, COLAMD, partial pivoting) on a {shape[0]}x{shape[1]}x{shape[2]} "
                  f"grid of the same generator: n={n}, factor {t_factor:.2f} s, {t_solve * 1e3:.1f} ms/solve; "
                  f"scaled by the factor-flop ratio {scale:.1f}x to the workload"
                  + (f"; like-for-like (the GPU path's nested-dissection order, no pivoting): factor {nd['sample_factor_s_nd_order_no_pivoting']:.2f} s"
                     if nd else "") if scale else "flop model unavailable",
        "sample_factor_s": t_factor, "sample_solve_s": t_solve, "sample_n": n, "host_cores": os.cpu_count(),
        "solves_per_sec_sample": 1.0 / t_solve,
    }


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    # algorithmic flops of the full workload (analysis only, CPU)
    full = build_case(wl)
    import ctypes
    sim = ctypes.CDLL(os.path.join(ROOT, "oracle", "libnkp_sim.so"))
    P = ctypes.POINTER
    stats = np.zeros(8)
    ip = lambda a: np.ascontiguousarray(a, dtype=np.int32).ctypes.data_as(P(ctypes.c_int))
    arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (full["rowptr"], full["colind"], *full["coords"])]
    sim.nkp_sim_run(full["n"], *(a.ctypes.data_as(P(ctypes.c_int)) for a in arrs[:2]),
                    full["nzval"].ctypes.data_as(P(ctypes.c_double)),
                    *(a.ctypes.data_as(P(ctypes.c_int)) for a in arrs[2:]), 64, 96, None, 0, None,
                    stats.ctypes.data_as(P(ctypes.c_double)), None, 1)
    full_flops = float(stats[5])
    vals = []
    cb = None
    for it in range(args.warmup + args.steps):
        cb = cpu_baseline(full_flops)
        if it >= args.warmup:
            vals.append(cb["value"])
        if it == 0 and cb["sample_factor_s"] * (args.warmup + args.steps) > 240:
            break  # keep the whole run within a few minutes
    v = float(np.mean(vals)) if vals else cb["value"]
    cb["value"] = v
    imt, jmt, km, desc = WORKLOADS[wl]
    line = {
        "impl": "reference", "metric": "numeric_factor_time_s", "value": v, "unit": "s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{desc}, n={full['n']}, nnz={len(full['nzval'])}, nrhs={NRHS}"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_own(args):
    import torch
    from nk_ocn_tracer_jacobian_precond_b200 import solver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    wl = args.workload
    imt, jmt, km, desc = WORKLOADS[wl]
    t0 = time.perf_counter()
    case = build_case(wl)
    t_gen = time.perf_counter() - t0
    n, nnz = case["n"], len(case["nzval"])
    comm = None
    if world > 1:
        # subtree-to-GPU sharding: NCCL communicator of the solver, id shipped through torch.distributed
        uid = [solver.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = (rank, world, uid[0])
    s = solver.TracerJacobianSolver(n, case["rowptr"], case["colind"], coords=case["coords"], comm=comm, device=local)
    st0 = s.stats()

    nsteps = args.warmup + args.steps
    rng = np.random.default_rng(1234)   # identical operands on every rank (A, B, X are replicated)
    import scipy.sparse as sp
    # Newton sequence: new values every step, same pattern (BASELINE.json configs[4])
    host_vals = [case["nzval"] * (1.0 + 1e-3 * rng.standard_normal(nnz)) for _ in range(min(nsteps, 3))]
    xs = rng.standard_normal((n, NRHS))
    host_B = []
    for v in host_vals:
        A = sp.csr_matrix((v, case["colind"], case["rowptr"]), shape=(n, n))
        host_B.append(np.asfortranarray(A @ xs))
    dev_vals = [torch.tensor(v, device=dev) for v in host_vals]
    dev_B = [torch.tensor(np.ascontiguousarray(b.T), device=dev) for b in host_B]   # (nrhs, n) row-major == column-major n x nrhs
    work_B = torch.empty_like(dev_B[0])

    def barrier():
        s.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---------------- device-resident arm -------------------------------------------------
    s.set_profile(True)
    fact_t, solve_t, gemm_t, refine = [], [], [], []
    launches0 = None
    sampler = None
    for it in range(nsteps):
        k = it % len(dev_vals)
        if it == args.warmup:
            barrier()
            launches0 = s.stats()["kernel_launches"]
            if rank == 0:
                sampler = ClockSampler(local)
                sampler.start()
            t_wall0 = time.perf_counter()
        s.factor_device(dev_vals[k].data_ptr())
        work_B.copy_(dev_B[k])
        torch.cuda.current_stream().synchronize()
        if dist is not None:
            dist.barrier()   # ranks leave the factorisation at different times; keep that skew out of the solve timing
        berr = s.solve_device(work_B.data_ptr(), n, NRHS)
        if it >= args.warmup:
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
    # correctness of the last step (manufactured solution)
    X = work_B.cpu().numpy().T
    k = (nsteps - 1) % len(dev_vals)
    A = sp.csr_matrix((host_vals[k], case["colind"], case["rowptr"]), shape=(n, n))
    relres = float((np.linalg.norm(A @ X - host_B[k], axis=0) / np.linalg.norm(host_B[k], axis=0)).max())
    solerr = float((np.linalg.norm(X - xs, axis=0) / np.linalg.norm(xs, axis=0)).max())

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
        work_B.copy_(dev_B[k])
        torch.cuda.current_stream().synchronize()
        if dist is not None:
            dist.barrier()
        s.solve_device(work_B.data_ptr(), n, NRHS)
        if it >= 2:
            nw_t.append(s.stats()["t_solve"])
            nw_steps.append(s.stats()["refine_steps"])
    Xn = work_B.cpu().numpy().T
    nw_relres = float((np.linalg.norm(A @ Xn - host_B[k], axis=0) / np.linalg.norm(host_B[k], axis=0)).max())
    nw_solerr = float((np.linalg.norm(Xn - xs, axis=0) / np.linalg.norm(xs, axis=0)).max())
    s.set_refine_rule(0)

    # residual SpMV r = b - A x (refinement): device time with torch events around the library call
    dx = torch.randn(NRHS, n, dtype=torch.float64, device=dev)
    dr = torch.empty_like(dx)
    spmv_t = []
    for it in range(3 + 5):
        s.sync()
        t0 = time.perf_counter()
        s.residual_device(dx.data_ptr(), dev_B[0].data_ptr(), dr.data_ptr(), NRHS)
        s.sync()
        if it >= 3:
            spmv_t.append(time.perf_counter() - t0)

    # ---------------- end-to-end arm (host buffers through the C ABI) ------------------------
    e2e_f, e2e_s = [], []
    Bh = np.empty_like(host_B[0])
    for it in range(min(args.warmup, 2) + args.steps):
        k = it % len(host_vals)
        Bh[:] = host_B[k]
        barrier()
        t0 = time.perf_counter()
        s.factor(host_vals[k])
        s.sync()
        if dist is not None:
            dist.barrier()
        t1 = time.perf_counter()
        s.solve(Bh)
        t2 = time.perf_counter()
        if it >= min(args.warmup, 2):
            e2e_f.append(t1 - t0)
            e2e_s.append(t2 - t1)

    def mx(v):
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    factor_s = mx(np.mean(fact_t))
    solve_s = mx(np.mean(solve_t))
    step_ms = mx(t_wall / args.steps * 1e3)
    e2e_factor = mx(np.mean(e2e_f))
    e2e_solve = mx(np.mean(e2e_s))
    sweep_s = mx(np.mean(sweep_t))
    nw_solve_s = mx(np.mean(nw_t))

    if rank == 0:
        peaks = measured_peaks()
        g = np.array(gemm_t)
        t_gemm, n_gemm = float(g[:, 0].mean()), float(g[:, 1].mean())
        gemm_tflops = st["gemm_flops"] / t_gemm * 1e-12 if t_gemm > 0 else None
        roofline = {
            "kernel": "k_gemm (Schur-complement update, FP64 DMMA)", "bound": "tensor",
            "achieved": gemm_tflops, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s",
            "frac": gemm_tflops / peaks["fp64_tflops"] if gemm_tflops else None,
            "peak_source": peaks["_fp64_src"], "traffic": None,
            "launches_per_factor": n_gemm, "avg_launch_ms": t_gemm / n_gemm * 1e3 if n_gemm else None,
            "alg_flops_per_launch": st["gemm_flops"] / n_gemm if n_gemm else None,
            "share_of_factor_time": t_gemm / float(np.mean(fact_t)),
            "factor_breakdown_s": {"gemm": t_gemm, "trsm": float(g[:, 2].mean()), "diag": float(g[:, 3].mean()),
                                   "extend_add": float(g[:, 4].mean()), "zero+scatter": float(g[:, 5].mean())},
            "factor_overall_tflops": st["factor_flops"] / factor_s * 1e-12,
        }
        solve_gbs = st["solve_bytes"] / sweep_s * 1e-9
        roofline_solve = {
            "kernel": "k_sweep_big + k_fwd_small/k_bwd_small sweep pair (nrhs=%d, no refinement)" % NRHS, "bound": "hbm",
            "achieved": solve_gbs, "peak": peaks["hbm_gbs"] * world, "unit": "GB/s",
            "frac": solve_gbs / (peaks["hbm_gbs"] * world),
            "peak_source": peaks["_hbm_src"] + (f" x {world} GPUs (whole-job bytes over the max-over-ranks time)" if world > 1 else ""),
            "traffic": None, "sweep_pair_ms": sweep_s * 1e3,
            "alg_bytes_per_sweep_pair": st["solve_bytes"],
        }
        spmv_s = float(np.min(spmv_t))
        spmv_bytes = NRHS * (12.0 * nnz + 4.0 * (n + 1) + 24.0 * n)   # BASELINE.md section 4, per right-hand side
        roofline_spmv = {
            "kernel": "k_residual (r = b - A x, %d rhs; host-timed launch+sync)" % NRHS, "bound": "hbm",
            "achieved": spmv_bytes / spmv_s * 1e-9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": spmv_bytes / spmv_s * 1e-9 / peaks["hbm_gbs"], "ms": spmv_s * 1e3, "traffic": None,
        }
        # DRAM traffic per launch from the committed ncu capture of this workload (null if none)
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath) and world == 1:
            tj = json.load(open(tpath))
            if wl in tj:
                roofline["traffic"] = tj[wl].get("k_gemm_bytes_per_launch")
                roofline["traffic_source"] = tj["source"]
                roofline_solve["traffic"] = tj[wl].get("sweep_pair_bytes")
                roofline_spmv["traffic"] = tj[wl].get("k_residual_bytes_per_launch")
        cb = cpu_baseline(st["factor_flops"])
        line = {
            "metric": "numeric_factor_time_s", "value": factor_s, "unit": "s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": False,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{desc}: one numeric refactorisation + one batched solve of {NRHS} tracers per step",
                "n": n, "nnz": nnz, "nrhs": NRHS, "options": "adv centered, hmix const, vmix const, sink const_shallow 365 10e2",
                "nnz_lu": st["nnz_lu"], "factor_flops": st["factor_flops"], "fronts": st["n_fronts"],
                "levels": st["n_levels"], "max_front": st["max_front"], "analysis_s": st0["t_analysis"],
                "l2": "operand and factors exceed L2 (%.1f GB heap); no flush needed" % (st["heap_bytes"] * 1e-9),
                "parallelism": "1 GPU" if world == 1 else
                f"{world} GPUs: nested-dissection subtrees sharded over the GPUs, top separator fronts on their heaviest "
                f"child's GPU, NCCL send/recv of update matrices and broadcast of separator solutions ({int(st['n_xfers'])} transfers)",
                "generate_s": t_gen,
            },
            "solve_s": solve_s, "solves_per_sec": NRHS / solve_s, "refine_steps": float(np.mean(refine)),
            "relres_max": relres, "solution_err_max": solerr, "berr_max": float(np.max(berr)),
            "tiny_pivots_replaced": int(st["tiny_pivots"]),
            "solve_normwise_rule": {"solve_s": nw_solve_s, "solves_per_sec": NRHS / nw_solve_s,
                                    "refine_steps": float(np.mean(nw_steps)), "relres_max": nw_relres,
                                    "solution_err_max": nw_solerr,
                                    "rule": "stop when ||b - A x||_2 <= 1e-14 ||b||_2 (nkp_options.refine_rule = 1)"},
            "roofline": roofline, "roofline_solve": roofline_solve, "roofline_spmv": roofline_spmv, "cpu_baseline": cb,
            "e2e": {"value": e2e_factor, "unit": "s", "h2d_bytes_per_step": 8 * nnz + 8 * n * NRHS,
                    "d2h_bytes_per_step": 8 * n * NRHS, "solve_s": e2e_solve, "solves_per_sec": NRHS / e2e_solve},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    s.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    faulthandler.enable()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
