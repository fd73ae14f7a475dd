set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s10_pytest.log
tail -4 gpurun_out/s10_pytest.log
for v in 1 0; do
  TT_DEFER_METRICS=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-tables --no-cpu-baseline --no-device-store 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('defer=$v', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step']*1e3,1),'us  e2e', round(d['e2e']['value']/1e6,3), 'pipelined', round(d['e2e_pipelined']['value']/1e6,3), 'loss', d['loss_last'])"
done
