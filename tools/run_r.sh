mkdir -p gpurun_out/r2
timeout 400 python -m pytest tests/test_gpu_long_keys.py tests/test_gpu_agg_fe.py -q --timeout 120 -x 2>&1 | tail -2
timeout 400 python -m pytest tests/test_gpu_parity.py -q --timeout 200 -x 2>&1 | tail -2
bash tools/fe_sweep.sh cfg3 default 6,7,3 4,7,4 6,8,2 | tee gpurun_out/r2/sweep3p.log
bash tools/fe_sweep.sh cfg5 default 6,7,3 4,7,4 | tee gpurun_out/r2/sweep5p.log
