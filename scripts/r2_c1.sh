mkdir -p gpurun_out
(cd scripts/microbench && ./mufu2 > ../../gpurun_out/c1_mufu2.log 2>&1); cat gpurun_out/c1_mufu2.log
timeout 900 python -m pytest tests/test_gpu_tower_fused.py -q > gpurun_out/c1_pytest_fused.log 2>&1; echo "pytest fused rc=$?"; tail -30 gpurun_out/c1_pytest_fused.log | cut -c1-300
