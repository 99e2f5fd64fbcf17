mkdir -p gpurun_out/r2d
timeout 100 python scripts/tree_ab.py 100 116 60 1:0.15:48,1:0.2:64,1:0.3:96 > gpurun_out/r2d/tree_ab_gx3v7_relax.log 2>&1; echo "rc=$?" >> gpurun_out/r2d/tree_ab_gx3v7_relax.log
cat gpurun_out/r2d/tree_ab_gx3v7_relax.log
timeout 200 python scripts/tree_ab.py 320 384 60 1:0.1:32,1:0.2:64 > gpurun_out/r2d/tree_ab_gx1v6.log 2>&1; echo "rc=$?" >> gpurun_out/r2d/tree_ab_gx1v6.log
cat gpurun_out/r2d/tree_ab_gx1v6.log
