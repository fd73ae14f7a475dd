# A/B of the built library against a baseline copy (ab/libtwotower_b200_base.so, built from HEAD before the change) on one box:
# full GPU test suite with the new build, stand-alone lookup timing, then interleaved short bench runs of both builds.
mkdir -p gpurun_out
T=${1:-p1}
LIB=jodalrob-twotower_b200/csrc/libtwotower_b200.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log | cut -c1-250
timeout 300 python scripts/time_lookup.py 2>&1 | grep -v Warn
cp $LIB /tmp/new.so
run() { timeout 600 python bench.py --steps 100 --warmup 10 --no-tables --no-cpu-baseline --no-device-store --no-parity 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$1', round(d['ms_per_step']*1e3,1), 'us', round(d['value']/1e6,2), 'M/s e2e', round(d['e2e']['value']/1e6,2), d['loss_last'])"; }
for i in 1 2; do
cp /tmp/new.so $LIB; run new
cp ab/libtwotower_b200_base.so $LIB; run base
done
cp /tmp/new.so $LIB
timeout 300 python scripts/trace_step.py > gpurun_out/${T}_trace.log 2>&1; echo "trace rc=$?"
