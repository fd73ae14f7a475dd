mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "sharded_multirank" > gpurun_out/i1_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/i1_pytest.log | cut -c1-300
TT_A2A=0 timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "sharded_multirank" > gpurun_out/i1_pytest_noa2a.log 2>&1; echo "pytest (TT_A2A=0) rc=$?"; tail -5 gpurun_out/i1_pytest_noa2a.log | cut -c1-300
