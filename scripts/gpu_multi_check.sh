# Multi-GPU check under `gpurun --gpus N`: golden-trajectory parity of the sharded step, the bench line, rank 0's step timeline.
set -x
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py > gpurun_out/m_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"
grep "dist_check\]" gpurun_out/m_dist_check_n$N.log | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/m_bench_n$N.json 2> gpurun_out/m_bench_n$N.err; echo "bench rc=$?"
tail -2 gpurun_out/m_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/trace_step_dist.py > gpurun_out/m_trace_n$N.log 2>&1
