# Round-2 profile pass on one B200 (under gpurun): the bench line, its ncu launch list, and `ncu --set full` captures of the dominant
# kernels (each command first runs without ncu). Outputs in gpurun_out/; summaries are copied to profiles/ by scripts/r2_collect_profiles.sh
T=${1:-p2}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 10 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tables --no-device-store --no-parity > /dev/null 2>&1; echo "short bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tables --no-device-store --no-parity > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
timeout 300 python scripts/time_simce3.py > gpurun_out/${T}_time_simce3.log 2>&1; echo "time_simce3 rc=$?"
for P in p1 p2 p3; do
TT_ONLY=$P timeout 600 ncu --set full --clock-control none --import-source on -k regex:simce_tc_kernel -s 3 -c 1 -o gpurun_out/${T}_simce_$P python scripts/time_simce3.py > gpurun_out/${T}_ncu_$P.log 2>&1; echo "ncu $P rc=$?"
done
timeout 300 python scripts/time_tower_fused.py > gpurun_out/${T}_time_tower.log 2>&1; echo "time_tower rc=$?"; cat gpurun_out/${T}_time_tower.log
for Kn in tower_tail_fwd_kernel tower_head_bwd_kernel tower_mid_bwd_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$Kn -s 3 -c 1 -o gpurun_out/${T}_$Kn python scripts/time_tower_fused.py > gpurun_out/${T}_ncu_$Kn.log 2>&1; echo "ncu $Kn rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 30 -c 10 -o gpurun_out/${T}_gemm_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tables --no-device-store --no-parity --no-graph > gpurun_out/${T}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 300 python scripts/trace_step.py > gpurun_out/${T}_trace.log 2>&1; echo "trace rc=$?"
ls -la gpurun_out | grep ${T}_ | awk '{print $5, $9}'
