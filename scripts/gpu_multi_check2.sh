set -x
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py > gpurun_out/m2_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"
grep "dist_check\]" gpurun_out/m2_dist_check_n$N.log | tail -4; tail -3 gpurun_out/m2_dist_check_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/m2_bench_n$N.json 2> gpurun_out/m2_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/m2_bench_n$N.err
python - <<PY
import json
for line in open('gpurun_out/m2_bench_n$N.json'):
    if line.startswith('{'):
        d=json.loads(line)
        for k in ("value","ms_per_step","e2e","e2e_device_store"): print(k, d.get(k))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/trace_step_dist.py > gpurun_out/m2_trace_n$N.log 2>&1
