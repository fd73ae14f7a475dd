set -x
timeout 600 python -m pytest tests -m gpu -q -x -k "sharded or fused or determin or golden or graphed" 2>&1 | tail -3
timeout 300 python bench.py --steps 100 --warmup 10 --no-tables --no-cpu-baseline --no-device-store 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step']*1e3,1),'us')"
bash scripts/gpu_multi_check2.sh 2
