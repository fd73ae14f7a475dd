set -x
timeout 900 python -m pytest tests/test_gpu_tables.py tests/test_gpu_parity.py -m gpu -q -k "tables or sparse or sort or golden" > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s5_pytest.log
tail -4 gpurun_out/s5_pytest.log
timeout 300 python bench.py --tables-only quick > gpurun_out/s5_tables.json 2>&1
TT_EBC_DIRECT=1 timeout 300 python bench.py --tables-only quick > gpurun_out/s5_tables_direct1.json 2>&1
cat gpurun_out/s5_tables.json gpurun_out/s5_tables_direct1.json
