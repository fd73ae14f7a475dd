set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s6_pytest.log
tail -6 gpurun_out/s6_pytest.log
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/s6_bench_n1.json 2> gpurun_out/s6_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/s6_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s6_bench_n1.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","e2e_device_store","roofline"): print(k, d.get(k))
print(json.dumps(d["tables"])[:1500])
PY
