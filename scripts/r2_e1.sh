mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_path.py -q -x -k "lazy" > gpurun_out/e1_pytest.log 2>&1; echo "pytest lazy rc=$?"; tail -25 gpurun_out/e1_pytest.log | cut -c1-250
run() { timeout 600 python bench.py --steps 100 --warmup 10 --no-tables --no-cpu-baseline --no-device-store --no-parity "$@" 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$*', round(d['ms_per_step']*1e3,1), 'us', round(d['value']/1e6,2), 'M/s', 'loss', d['loss_last'])"; }
run --optimizer dense_exact
run --optimizer lazy_exact
run --optimizer sparse
