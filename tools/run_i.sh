mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2/gpu_all3.log 2>&1; tail -4 gpurun_out/r2/gpu_all3.log
for rows in 1000000 8000000 64000000; do
  timeout 120 python bench.py --workload cfg3 --rows $rows --steps 20 --no-sub --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c 'import sys,json; l=json.loads(sys.stdin.read()); print(l["config"]["rows_total"], "ms/step", l["ms_per_step"], "launches", l["gpu_launches"])' | tee -a gpurun_out/r2/rowscale2.log
done
bash tools/fe_sweep.sh cfg3 default 8,7,2 > gpurun_out/r2/sweep3k.log 2>&1; cat gpurun_out/r2/sweep3k.log
bash tools/fe_sweep.sh cfg5 default > gpurun_out/r2/sweep5k.log 2>&1; cat gpurun_out/r2/sweep5k.log
bash tools/fe_sweep.sh cfg2f default > gpurun_out/r2/sweep2k.log 2>&1; cat gpurun_out/r2/sweep2k.log
bash tools/fe_sweep.sh cfg4 default > gpurun_out/r2/sweep4k.log 2>&1; cat gpurun_out/r2/sweep4k.log
