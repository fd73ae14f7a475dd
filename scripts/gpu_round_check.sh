set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
tail -5 gpurun_out/s2_pytest.log
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/s2_bench_n1.json 2> gpurun_out/s2_bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/s2_bench_n1.json
TT_EBC_DIRECT=0 timeout 300 python bench.py --tables-only quick > gpurun_out/s2_tables_direct0.json 2>&1
TT_EBC_DIRECT=1 timeout 300 python bench.py --tables-only quick > gpurun_out/s2_tables_direct1.json 2>&1
cat gpurun_out/s2_tables_direct0.json gpurun_out/s2_tables_direct1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tables > gpurun_out/s2_ncu1.log 2>&1; echo "ncu1 rc=$?"
TT_ONLY=bwd timeout 600 ncu --set full --clock-control none --import-source on -k regex:simce_tc_kernel -s 2 -c 2 -o gpurun_out/s2_simce_bwd python scripts/time_simce.py > gpurun_out/s2_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 600 python scripts/time_tables.py > gpurun_out/s2_time_tables.log 2>&1; cat gpurun_out/s2_time_tables.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ebc_fwd_jagged|ebc_rows_sparse" -s 2 -c 2 -o gpurun_out/s2_tables python scripts/time_tables.py > gpurun_out/s2_ncu3.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out
