mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_agg_fe.py tests/test_gpu_long_keys.py -q --timeout 120 -x > gpurun_out/r2/fe_a.log 2>&1; tail -3 gpurun_out/r2/fe_a.log
bash tools/fe_sweep.sh cfg3 default 8,7,2 8,6,2 > gpurun_out/r2/sweep3g.log 2>&1; cat gpurun_out/r2/sweep3g.log
bash tools/fe_sweep.sh cfg5 default 8,6,2 > gpurun_out/r2/sweep5g.log 2>&1; cat gpurun_out/r2/sweep5g.log
bash tools/ncu_agg.sh cfg3 400000000 fe3g
