mkdir -p gpurun_out/r2
bash tools/ncu_agg.sh cfg3 400000000 fe3j
KQ_TIME_AGG=1 timeout 120 python tools/step_cost.py 1 1000000 cfg3 2> gpurun_out/r2/timeagg.err | tail -2 | tee gpurun_out/r2/stepcost2.log
tail -4 gpurun_out/r2/timeagg.err
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r2/gpu_all.log 2>&1; tail -5 gpurun_out/r2/gpu_all.log
