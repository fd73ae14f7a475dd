mkdir -p gpurun_out
TT_KINETO=1 timeout 300 python scripts/time_simce3.py 2>/dev/null | head -8 > gpurun_out/c2_time3_poly.log; cat gpurun_out/c2_time3_poly.log
TT_KNOBS=3 timeout 300 python scripts/time_simce3.py 2>/dev/null | head -8 > gpurun_out/c2_time3_nopoly.log; cat gpurun_out/c2_time3_nopoly.log
bash scripts/r2_check.sh c2
