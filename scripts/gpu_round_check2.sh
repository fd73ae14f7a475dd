set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
tail -8 gpurun_out/s3_pytest.log
timeout 300 python bench.py --tables-only full > gpurun_out/s3_tables.json 2>&1
TT_EBC_DIRECT=1 timeout 300 python bench.py --tables-only quick > gpurun_out/s3_tables_direct1.json 2>&1
cat gpurun_out/s3_tables.json gpurun_out/s3_tables_direct1.json
timeout 300 python scripts/time_tables.py > gpurun_out/s3_time_tables.log 2>&1; cat gpurun_out/s3_time_tables.log
