mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "multihot or emulated or sharded" > gpurun_out/h1_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/h1_pytest.log | cut -c1-250
