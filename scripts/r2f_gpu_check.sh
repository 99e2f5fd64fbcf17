mkdir -p gpurun_out/r2f
timeout 120 python -m pytest tests -x -q -m gpu > gpurun_out/r2f/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f/pytest_gpu.log
tail -4 gpurun_out/r2f/pytest_gpu.log
timeout 150 python bench.py --steps 3 --warmup 3 > gpurun_out/r2f/bench_default.json 2> gpurun_out/r2f/bench_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2f/bench_default.err; head -c 1500 gpurun_out/r2f/bench_default.json
