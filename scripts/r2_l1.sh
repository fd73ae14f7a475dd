mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_path.py -x -q -k "compact or graph or standalone" > gpurun_out/l1_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/l1_pytest.log | cut -c1-250
timeout 900 python bench.py --steps 100 --warmup 10 --no-tables --no-cpu-baseline > gpurun_out/l1_bench.json 2> gpurun_out/l1_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/l1_bench.err
python - <<PY
import json
for ln in open("gpurun_out/l1_bench.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print(round(d["ms_per_step"]*1e3,1), "us", round(d["value"]/1e6,2), "M/s; e2e", round(d["e2e"]["value"]/1e6,2), "pipelined", round(d["e2e_pipelined"]["value"]/1e6,2), "store", round(d["e2e_device_store"]["value"]/1e6,2), d["e2e"]["h2d_bytes_per_step"], d["launches_per_step"], d["parity"]["ok"])
PY
