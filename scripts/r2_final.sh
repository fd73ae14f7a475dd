# Final single-GPU pass of round 2: tests, smoke, the default bench line, the reference arm, ncu launch list, stand-alone kernel
# timings and an `ncu --set full` capture of the rewritten lookup. Outputs in gpurun_out/ (tag f1).
T=${1:-f1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${T}_smoke.log
timeout 1200 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench_n1.err
python - <<PY
import json
try:
    for ln in open("gpurun_out/${T}_bench_n1.json"):
        if ln.startswith("{"):
            d = json.loads(ln)
            print({k: d[k] for k in ("value", "ms_per_step", "launches_per_step", "parity", "loss_last", "clocks") if k in d})
            print("e2e", d["e2e"], "\nroofline", d["roofline"], "\ncpu", d["cpu_baseline"], "\nhbm", d.get("hbm"))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "reference arm rc=$?"; cat gpurun_out/${T}_bench_ref.json | cut -c1-400
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tables --no-device-store --no-parity > /dev/null 2>&1; echo "short bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tables --no-device-store --no-parity > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
timeout 300 python scripts/time_simce3.py > gpurun_out/${T}_time_simce3.log 2>&1; echo "time_simce3 rc=$?"
timeout 300 python scripts/time_tower_fused.py > gpurun_out/${T}_time_tower.log 2>&1; echo "time_tower rc=$?"
timeout 300 python scripts/time_lookup.py > gpurun_out/${T}_time_lookup.log 2>&1; echo "time_lookup rc=$?"; grep -v Warn gpurun_out/${T}_time_lookup.log
TT_ONLY=notice timeout 600 ncu --set full --clock-control none --import-source on -k regex:ebc_fwd_l1 -s 4 -c 1 -o gpurun_out/${T}_lookup python scripts/time_lookup.py > gpurun_out/${T}_ncu_lookup.log 2>&1; echo "ncu lookup rc=$?"
timeout 300 python scripts/trace_step.py > gpurun_out/${T}_trace.log 2>&1; echo "trace rc=$?"; head -3 gpurun_out/${T}_trace.log | tail -1
