mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_long_keys.py tests/test_gpu_agg_fe.py -q --timeout 120 -x 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 200 -x 2>&1 | tail -2
one() { timeout 120 python bench.py --workload $2 --rows $1 --steps 20 --no-sub --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c 'import sys,json; l=json.loads(sys.stdin.read()); print("ms/step", l["ms_per_step"], l["check"].get("count_ok"))'; }
echo "cfg3 1M: $(one 1000000 cfg3)"
echo "cfg3 125M: $(one 125000000 cfg3)"
echo "cfg5 1M: $(one 1000000 cfg5)"
KQ_TIME_AGG=1 timeout 120 python tools/step_cost.py 1 1000000 cfg3 2>&1 | tail -2
bash tools/fe_sweep.sh cfg3 default 8,7,2 | tee gpurun_out/r2/sweep3l.log
bash tools/fe_sweep.sh cfg5 default | tee gpurun_out/r2/sweep5l.log
