mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_agg_fe.py tests/test_gpu_long_keys.py -q --timeout 120 -x > gpurun_out/r2/fe_b.log 2>&1; tail -3 gpurun_out/r2/fe_b.log
bash tools/fe_sweep.sh cfg3 default 8,7,2 8,3,3,2 8,3,2,2 4,3,4,2 4,3,6,2 6,3,3,2 > gpurun_out/r2/sweep3h.log 2>&1; cat gpurun_out/r2/sweep3h.log
bash tools/fe_sweep.sh cfg5 default 8,6,2 8,3,2,2 4,3,3,2 4,3,4,2 > gpurun_out/r2/sweep5h.log 2>&1; cat gpurun_out/r2/sweep5h.log
