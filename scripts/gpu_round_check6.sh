set -x
timeout 900 python -m pytest tests -m gpu -q -k "sharded or loader or eval" > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_pytest.log
tail -6 gpurun_out/s7_pytest.log
TT_ONLY=fwd timeout 600 ncu --set full --clock-control none --import-source on -k regex:simce_tc_kernel -s 2 -c 2 -o gpurun_out/s7_simce_fwd python scripts/time_simce.py > gpurun_out/s7_ncu_fwd.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ebc_rows_sparse|ebc_piece" -s 2 -c 2 -o gpurun_out/s7_sparse python scripts/time_tables.py > gpurun_out/s7_ncu_sparse.log 2>&1; echo "ncu rc=$?"
python scripts/time_simce.py > gpurun_out/s7_time_simce.log 2>&1; cat gpurun_out/s7_time_simce.log
