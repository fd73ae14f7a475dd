# Turn the gpurun_out/ artefacts of scripts/r2_profiles.sh (tag $1) into the tracked summaries under profiles/ (round 2).
T=${1:-p2}
O=gpurun_out
python scripts/summarize_ncu_launches.py $O/${T}_launches.csv "ncu launch list of \`python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tables --no-device-store --no-parity\` (round 2)" > profiles/r02_ncu_launch_list_summary.md
cp $O/${T}_launches.csv profiles/r02_ncu_launches.csv
for P in p1 p2 p3; do
python scripts/summarize_ncu_report.py $O/${T}_simce_$P.ncu-rep profiles/r02_ncu_simce_$P.json "TT_ONLY=$P ncu --set full --clock-control none --import-source on -k regex:simce_tc_kernel -s 3 -c 1 python scripts/time_simce3.py" "three-pass logits step, pass $P, B=8192 x 8192, F=64, T=1 (README-default)" > /dev/null
done
for K in tower_tail_fwd_kernel tower_head_bwd_kernel tower_mid_bwd_kernel; do
python scripts/summarize_ncu_report.py $O/${T}_$K.ncu-rep profiles/r02_ncu_$K.json "ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 python scripts/time_tower_fused.py" "fused tower kernel, notice tower shapes (M=8192, H1=64, F=64, LD=1152)" > /dev/null
done
python scripts/summarize_ncu_report.py $O/${T}_gemm_tc.ncu-rep profiles/r02_ncu_gemm_tc.json "ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 30 -c 10 python bench.py --steps 2 --warmup 3 --no-graph ..." "ten consecutive gemm_tc_kernel launches of an eager step: forward (K-major), weight-gradient (MN-major, split-K) variants" > /dev/null
python - <<PY
import json
for ln in open("$O/${T}_bench_n1.json"):
    if ln.startswith("{"):
        json.dump(json.loads(ln), open("profiles/r02_bench_n1.json", "w"), indent=1)
PY
cp $O/${T}_trace.log profiles/r02_trace_n1.log
cp $O/${T}_time_simce3.log profiles/r02_time_simce3.log
cp $O/${T}_time_tower.log profiles/r02_time_tower_fused.log
ls -la profiles | grep r02
