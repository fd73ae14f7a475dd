# One 8-GPU session: golden / bf16 parity of the sharded step, the headline bench line at N=8, BASELINE configs[2] and configs[4]
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py > gpurun_out/n${N}_dist_check.log 2>&1; echo "dist_check rc=$?"
grep "dist_check\]" gpurun_out/n${N}_dist_check.log | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/n${N}_bench.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print("N=$N", round(d["ms_per_step"]*1e3,1), "us", round(d["value"]/1e6,2), "M/s", d["parity"]["ok"], d["scaling_model"])
PY
bash scripts/r2_cfg3.sh $N 10000000 lazy_exact sparse dense_exact
CFG=cfg5 bash scripts/r2_cfg3.sh $N 0 sparse dense_exact
