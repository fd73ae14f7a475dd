# A/B of an engine env knob on one box: full GPU test suite, then interleaved short bench runs with the knob off / on.
# usage: bash scripts/r2_abenv.sh <tag> <ENV_NAME> <value_a> <value_b>
mkdir -p gpurun_out
T=${1:-q1}; K=${2:-TT_SIMCE_FUSED_FINISH}; A=${3:-0}; B=${4:-1}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 300 python scripts/time_simce3.py 2>&1 | grep "simce_tc_kernel\|tc_finish\|tc_dx"
run() { env "$@" timeout 600 python bench.py --steps 100 --warmup 10 --no-tables --no-cpu-baseline --no-device-store --no-parity 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$*', round(d['ms_per_step']*1e3,1), 'us', round(d['value']/1e6,2), 'M/s e2e', round(d['e2e']['value']/1e6,2), d['launches_per_step'], d['loss_last'])"; }
for i in 1 2; do
run $K=$A
run $K=$B
done
timeout 300 python scripts/trace_step.py > gpurun_out/${T}_trace.log 2>&1; echo "trace rc=$?"
