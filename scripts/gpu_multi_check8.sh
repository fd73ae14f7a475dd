set -x
N=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py > gpurun_out/m8_dist_check.log 2>&1; echo "dist_check rc=$?"
grep "dist_check\]" gpurun_out/m8_dist_check.log | tail -3; tail -2 gpurun_out/m8_dist_check.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 50 --warmup 10 --no-device-store > gpurun_out/m8_bench.json 2> gpurun_out/m8_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/m8_bench.err
python - <<PY
import json
for line in open('gpurun_out/m8_bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        for k in ("value","ms_per_step","e2e"): print(k, d.get(k))
PY
