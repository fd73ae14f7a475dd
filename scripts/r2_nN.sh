# bench line (and, with a second argument, the GPU test suite first) on N GPUs of one box: bash scripts/r2_nN.sh N [tests]
N=${1:-2}
mkdir -p gpurun_out
if [ -n "$2" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/g${N}_pytest.log | cut -c1-300
else
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py > gpurun_out/g${N}_dist_check.log 2>&1; echo "dist_check rc=$?"
grep "dist_check\]" gpurun_out/g${N}_dist_check.log | tail -6
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/g${N}_bench.json 2> gpurun_out/g${N}_bench.err; echo "bench rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/g${N}_bench.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print("N=$N", round(d["ms_per_step"]*1e3,1), "us", round(d["value"]/1e6,2), "M/s e2e", round(d["e2e"]["value"]/1e6,2), d["parity"]["ok"], d["scaling_model"])
PY
