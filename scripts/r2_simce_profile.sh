mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "bench_path" > gpurun_out/b2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/b2_pytest.log
timeout 300 python scripts/time_simce3.py > gpurun_out/b2_time3.log 2>&1; echo "time3 rc=$?"
for P in p1 p2 p3; do
TT_ONLY=$P timeout 600 ncu --set full --clock-control none --import-source on -k regex:simce_tc_kernel -s 1 -c 1 -o gpurun_out/b2_simce_$P python scripts/time_simce3.py > gpurun_out/b2_ncu_$P.log 2>&1; echo "ncu $P rc=$?"
done
