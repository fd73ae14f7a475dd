# What a round-end check runs on one B200 (under gpurun): parity tests, the bench line, the ncu launch list of the same bench
# command, and the `ncu --set full` captures of the dominant kernels. Outputs land in gpurun_out/ (copy what is judged to profiles/).
set -x
T=${1:-final}
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tables --no-device-store > gpurun_out/${T}_ncu1.log 2>&1; echo "ncu launch list rc=$?"
TT_ONLY=fwd timeout 600 ncu --set full --clock-control none --import-source on -k regex:simce_tc_kernel -s 2 -c 2 -o gpurun_out/${T}_simce_fwd python scripts/time_simce.py > gpurun_out/${T}_ncu2.log 2>&1; echo "ncu fwd rc=$?"
TT_ONLY=bwd timeout 600 ncu --set full --clock-control none --import-source on -k regex:simce_tc_kernel -s 2 -c 2 -o gpurun_out/${T}_simce_bwd python scripts/time_simce.py > gpurun_out/${T}_ncu3.log 2>&1; echo "ncu bwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ebc_fwd_jagged|ebc_rows_sparse" -s 2 -c 2 -o gpurun_out/${T}_tables python scripts/time_tables.py > gpurun_out/${T}_ncu4.log 2>&1; echo "ncu tables rc=$?"
python scripts/trace_step.py > gpurun_out/${T}_trace.log 2>&1
ls -la gpurun_out | tail -15
