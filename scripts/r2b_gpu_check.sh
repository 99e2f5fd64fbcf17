set -x
mkdir -p gpurun_out/r2b
timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -k "largediag or reference_driver_cli or golden_vectors" > gpurun_out/r2b/pytest_rowperm.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b/pytest_rowperm.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2b/smoke.log
timeout 120 python scripts/rowperm_probe.py > gpurun_out/r2b/rowperm_probe_gx3v7.log 2>&1; echo "probe rc=$?" | tee -a gpurun_out/r2b/rowperm_probe_gx3v7.log
tail -3 gpurun_out/r2b/pytest_rowperm.log; tail -2 gpurun_out/r2b/smoke.log; cat gpurun_out/r2b/rowperm_probe_gx3v7.log
