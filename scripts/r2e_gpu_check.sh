mkdir -p gpurun_out/r2e
timeout 100 python scripts/tree_ab.py 100 116 60 1 > gpurun_out/r2e/tree_ab_gx3v7.log 2>&1; echo "rc=$?" >> gpurun_out/r2e/tree_ab_gx3v7.log
cat gpurun_out/r2e/tree_ab_gx3v7.log
timeout 150 python scripts/tree_ab.py 320 384 60 1 > gpurun_out/r2e/tree_ab_gx1v6.log 2>&1; echo "rc=$?" >> gpurun_out/r2e/tree_ab_gx1v6.log
cat gpurun_out/r2e/tree_ab_gx1v6.log
