# BASELINE configs[2] under `gpurun --gpus N`: bash scripts/r2_cfg3.sh N [rows_per_table] [optimizers...]
N=${1:-8}; ROWS=${2:-10000000}; CFG=${CFG:-cfg3}; shift; shift
OPTS=${@:-lazy_exact sparse dense_exact}
mkdir -p gpurun_out
for O in $OPTS; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --config $CFG --rows-per-table $ROWS --optimizer $O --steps 20 --warmup 5 > gpurun_out/${CFG}_n${N}_$O.json 2> gpurun_out/${CFG}_n${N}_$O.err; echo "cfg3 N=$N $O rc=$?"
python - <<PY
import json
try:
    for ln in open("gpurun_out/${CFG}_n${N}_$O.json"):
        if ln.startswith("{"):
            d=json.loads(ln)
            print("$O", round(d["ms_per_step"],3), "ms", round(d["value"]/1e6,2), "M samples/s", d["ids_per_rank"], "nvlink MB", d["nvlink_total_MB_per_rank_per_step"])
            for r in d["kernels"][:10]: print("    ", r)
except Exception as e:
    print("parse failed", e)
PY
tail -3 gpurun_out/${CFG}_n${N}_$O.err
done
