N=4
mkdir -p gpurun_out
for A in 1 0; do
TT_A2A=$A timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$A tests/dist_check.py > gpurun_out/n${N}_dist_check_a2a$A.log 2>&1; echo "dist_check TT_A2A=$A rc=$?"
grep "dist_check\]" gpurun_out/n${N}_dist_check_a2a$A.log | tail -6
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/n${N}_bench.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print("N=$N", round(d["ms_per_step"]*1e3,1), "us", round(d["value"]/1e6,2), "M/s", d["parity"]["ok"], d["scaling_model"])
PY
