# Round-2 single-GPU check: parity tests, smoke, bench line, kernel timeline. Usage: bash scripts/r2_check.sh TAG [pytest-k-expr]
T=${1:-r2}
K=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
if [ -n "$K" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
else
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
fi
tail -15 gpurun_out/${T}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 50 --warmup 10 --no-tables > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench_n1.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_n1.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "launches_per_step", "parity", "loss_last") if k in d})
    print("e2e", d["e2e"]["value"], "roofline", d["roofline"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python scripts/trace_step.py > gpurun_out/${T}_trace.log 2>&1; echo "trace rc=$?"; head -3 gpurun_out/${T}_trace.log
