mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/nn1_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/nn1_pytest.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
run() { timeout 600 python bench.py --steps 100 --warmup 10 --no-tables --no-cpu-baseline --no-device-store --no-parity "$@" 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$*', round(d['ms_per_step']*1e3,1), 'us', round(d['value']/1e6,2), 'M/s', d['loss_last'])"; }
run --optimizer dense_exact
run --optimizer lazy_exact
