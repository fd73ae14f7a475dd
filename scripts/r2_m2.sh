run() { env "$@" timeout 600 python bench.py --steps 100 --warmup 10 --no-tables --no-cpu-baseline --no-device-store --no-parity 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$*', round(d['ms_per_step']*1e3,1), 'us', round(d['value']/1e6,2), 'M/s', d['loss_last'])"; }
run TT_CRIT_PRIO=0
run TT_CRIT_PRIO=1
run TT_CRIT_PRIO=0
run TT_CRIT_PRIO=1
