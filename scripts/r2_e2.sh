mkdir -p gpurun_out
for O in lazy_exact dense_exact; do
timeout 600 python bench.py --steps 100 --warmup 10 --no-tables --no-cpu-baseline --no-device-store --no-parity --optimizer $O 2>/dev/null > gpurun_out/e2_$O.json
python - <<PY
import json
for ln in open("gpurun_out/e2_$O.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print("$O", round(d["ms_per_step"]*1e3,1))
        for r in d["kernels"][:16]: print("   ", r["kernel"][:60], r["us_per_step"], r["launches_per_step"])
PY
done
