import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "solve_s", "solves_per_sec", "refine_steps", "relres_max", "solution_err_max", "ms_per_step", "gpu_launches")})
r = d["roofline"]; print("gemm TF/s", round(r["achieved"], 2), "frac", round(r["frac"], 3), "share", round(r["share_of_factor_time"], 3), "overall TF/s", round(r["factor_overall_tflops"], 2), r["factor_breakdown_s"])
r = d["roofline_solve"]; print("sweep GB/s", round(r["achieved"], 1), "frac", round(r["frac"], 3), "sweep_pair_ms", round(r["sweep_pair_ms"], 3))
print("e2e", d["e2e"])
if "roofline_spmv" in d:
    r = d["roofline_spmv"]; print("spmv GB/s", round(r["achieved"], 1), "frac", round(r["frac"], 3), "ms", round(r["ms"], 3))
