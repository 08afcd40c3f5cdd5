mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_long_keys.py tests/test_gpu_agg_fe.py -q --timeout 120 -x 2>&1 | tail -2
one() { timeout 120 python bench.py --workload cfg3 --rows $1 --steps 20 --no-sub --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c 'import sys,json; l=json.loads(sys.stdin.read()); print("ms/step", l["ms_per_step"])'; }
echo "base 1M: $(one 1000000)"
echo "noexact 1M: $(KQ_FE_NOEXACT=1 one 1000000)"
echo "nomerge 1M: $(KQ_FE_NOMERGE=1 one 1000000)"
echo "norefresh 1M: $(KQ_FE_NOREFRESH=1 one 1000000)"
echo "noexact+nomerge 1M: $(KQ_FE_NOEXACT=1 KQ_FE_NOMERGE=1 one 1000000)"
echo "geom 4,4 1M: $(KQ_AGG_GEOM=4,4 one 1000000)"
KQ_TIME_AGG=1 timeout 120 python tools/step_cost.py 1 1000000 cfg3 2>&1 | tail -3
