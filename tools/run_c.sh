mkdir -p gpurun_out/r2
bash tools/fe_sweep.sh cfg3 12,6,2 16,5,2 6,8,2 10,6,2 12,5,3 > gpurun_out/r2/sweep3i.log 2>&1; cat gpurun_out/r2/sweep3i.log
bash tools/fe_sweep.sh cfg5 12,6,2 16,4,2 10,6,2 > gpurun_out/r2/sweep5i.log 2>&1; cat gpurun_out/r2/sweep5i.log
timeout 120 python tools/step_cost.py 1 1000000 cfg3 2>&1 | tail -2 | tee gpurun_out/r2/stepcost1.log
timeout 120 python tools/step_cost.py 1 1000000 cfg4 2>&1 | tail -2 | tee -a gpurun_out/r2/stepcost1.log
