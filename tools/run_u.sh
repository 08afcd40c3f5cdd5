KQ_TRACE_AGG=1 timeout 100 python bench.py --workload cfg3 --steps 10 --no-sub --no-e2e --no-cpu-baseline 2>&1 | grep "kq fe geometry" | head -1
KQ_AGG_DIR=512 KQ_TRACE_AGG=1 timeout 100 python bench.py --workload cfg3 --steps 10 --no-sub --no-e2e --no-cpu-baseline 2>&1 | grep "kq fe geometry" | head -1
KQ_AGG_DIR=512 bash tools/fe_sweep.sh cfg3 default 8,7,3
KQ_AGG_DIR=256 bash tools/fe_sweep.sh cfg3 8,7,3
