mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tower_fused.py -q -x > gpurun_out/c3_pytest_fused.log 2>&1; echo "pytest fused rc=$?"; tail -3 gpurun_out/c3_pytest_fused.log
timeout 300 python scripts/time_tower_fused.py > gpurun_out/c3_time_tower.log 2>&1; cat gpurun_out/c3_time_tower.log
timeout 600 python bench.py --steps 50 --warmup 10 --no-tables --no-cpu-baseline --no-device-store > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c3_bench.json')); print(d['value'], d['ms_per_step'], d['launches_per_step'], d['parity']['ok'])"
timeout 300 python scripts/trace_step.py > gpurun_out/c3_trace.log 2>&1
