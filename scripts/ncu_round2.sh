#!/bin/bash
# Round-2 profiling pass (one GPU): plain run first, then the launch list and the --set full captures.
set -x
mkdir -p gpurun_out/r2prof
CMD="python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu"
$CMD > gpurun_out/r2prof/plain_bench.json 2> gpurun_out/r2prof/plain_bench.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2prof/launches_bench_gx1v6.csv $CMD > gpurun_out/r2prof/ncu_launches.log 2>&1
echo "launch list rc=$?"
# dominant kernel of the factorisation: one of the big wide-update launches
python scripts/sweep_ab.py gx1v6 0 > gpurun_out/r2prof/plain_sweep_ab.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm -s 600 -c 2 -o gpurun_out/r2prof/gemm python scripts/sweep_ab.py gx1v6 0 > gpurun_out/r2prof/ncu_gemm.log 2>&1
echo "gemm rc=$?"
# one sweep pair + residual (the factorisation's kernels are skipped by name)
timeout 700 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_big|k_fwd_front|k_bwd_front|k_fwd_small|k_bwd_small|k_residual|k_gather_bnd" -c 64 -o gpurun_out/r2prof/sweeps python scripts/sweep_ab.py gx1v6 0 > gpurun_out/r2prof/ncu_sweeps.log 2>&1
echo "sweeps rc=$?"
ls -la gpurun_out/r2prof
