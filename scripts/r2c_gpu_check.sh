mkdir -p gpurun_out/r2c
timeout 150 python scripts/tree_ab.py 100 116 60 0,1 > gpurun_out/r2c/tree_ab_gx3v7.log 2>&1; echo "ab rc=$?" >> gpurun_out/r2c/tree_ab_gx3v7.log
cat gpurun_out/r2c/tree_ab_gx3v7.log
timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/r2c/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest_gpu.log
tail -5 gpurun_out/r2c/pytest_gpu.log
