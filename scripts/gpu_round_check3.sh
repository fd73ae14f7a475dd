set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest.log
tail -5 gpurun_out/s4_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ebc_rows_sparse|ebc_gather_sorted" -s 2 -c 2 -o gpurun_out/s4_sparse python scripts/time_tables.py > gpurun_out/s4_ncu_sparse.log 2>&1; echo "ncu rc=$?"
TT_EBC_DIRECT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ebc_rows_sparse" -s 1 -c 1 -o gpurun_out/s4_sparse_direct python scripts/time_tables.py > gpurun_out/s4_ncu_sparse_d.log 2>&1; echo "ncu rc=$?"
TT_MODE=dense_exact timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ebc_rows_dense_stream|ebc_fwd_jagged" -s 2 -c 2 -o gpurun_out/s4_dense python scripts/time_tables.py > gpurun_out/s4_ncu_dense.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/s4_ncu_dense.log | tail -5
