mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pair_alignment or deterministic or fused_step_matches" > gpurun_out/m1_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/m1_pytest.log | cut -c1-250
run() { env "$@" timeout 600 python bench.py --steps 100 --warmup 10 --no-tables --no-cpu-baseline --no-device-store --no-parity 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$*', round(d['ms_per_step']*1e3,1), 'us', round(d['value']/1e6,2), 'M/s', d['loss_last'])"; }
run TT_SPLIT_ADAM=0
run TT_SPLIT_ADAM=2
run TT_SPLIT_ADAM=0
run TT_SPLIT_ADAM=2
