mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "emulated or all_to_all or sharded or lazy" > gpurun_out/f1_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/f1_pytest.log | cut -c1-250
