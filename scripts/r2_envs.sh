# interleaved short bench runs under different env settings on one box: bash scripts/r2_envs.sh "A=1" "A=2 B=3" ... (two rounds)
mkdir -p gpurun_out
run() { env $1 timeout 600 python bench.py --steps 100 --warmup 10 --no-tables --no-cpu-baseline --no-device-store --no-parity 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$1', round(d['ms_per_step']*1e3,1), 'us', round(d['value']/1e6,2), 'M/s e2e', round(d['e2e']['value']/1e6,2), d['launches_per_step'], d['loss_last'])"; }
for i in 1 2; do
for cfg in "$@"; do run "$cfg"; done
done
