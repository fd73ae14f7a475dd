set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s9_pytest.log
tail -4 gpurun_out/s9_pytest.log
for cfg in "1 1" "0 1"; do
  set -- $cfg
  TT_SPLIT_ADAM=$1 TT_HYPER_HELPER=$2 timeout 300 python bench.py --steps 100 --warmup 10 --no-tables --no-cpu-baseline --no-device-store 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('split=$1 hyper_helper=$2', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step']*1e3,1),'us  e2e', round(d['e2e']['value']/1e6,3), 'launches', d['launches_per_step'])"
done
python scripts/trace_step.py > gpurun_out/s9_trace.log 2>&1
