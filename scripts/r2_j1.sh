mkdir -p gpurun_out
for i in 1 2 3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$i tests/dist_check.py > gpurun_out/j1_dist_check_$i.log 2>&1; echo "dist_check run $i rc=$?"
done
grep "dist_check\]" gpurun_out/j1_dist_check_3.log | tail -6
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j1_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/j1_pytest.log
