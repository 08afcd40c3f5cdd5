mkdir -p gpurun_out/r2
timeout 200 python tools/part_dup.py 0.05 two > gpurun_out/r2/partdup.log 2>&1; tail -25 gpurun_out/r2/partdup.log | cut -c1-300
timeout 200 python tools/part_dup.py 0.05 one >> gpurun_out/r2/partdup.log 2>&1; tail -6 gpurun_out/r2/partdup.log | cut -c1-300
for rows in 1000000 8000000 64000000 250000000; do
  timeout 120 python bench.py --workload cfg3 --rows $rows --steps 20 --no-sub --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c 'import sys,json; l=json.loads(sys.stdin.read()); print(l["config"]["rows_total"], "ms/step", l["ms_per_step"], "launches", l["gpu_launches"])' | tee -a gpurun_out/r2/rowscale.log
done
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2/gpu_all2.log 2>&1; tail -8 gpurun_out/r2/gpu_all2.log
