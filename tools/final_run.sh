#!/bin/bash
# Round-end measurement on one B200: tests, smoke, the bench lines, the launch list and the ncu captures that profiles/ summarises.
O=gpurun_out/final; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err; cut -c1-300 $O/bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for w in cfg3 cfg4 cfg5 cfg2i csv; do python bench.py --workload $w --steps 5 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; cut -c1-120 $O/bench_$w.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg4.csv python bench.py --workload cfg4 --rows 100000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_csv.csv python bench.py --workload csv --rows 2000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches_csv.log 2>&1
N="ncu --set full --clock-control none --import-source on -c 1"
$N -k kq_hash_aggregate -s 3 -o $O/agg_cfg3 python tools/agg_one.py cfg3 100000000 > $O/ncu_cfg3.log 2>&1
$N -k kq_hash_aggregate -s 3 -o $O/agg_cfg5 python tools/agg_one.py cfg5 100000000 > $O/ncu_cfg5.log 2>&1
$N -k kq_hash_aggregate -s 3 -o $O/agg_cfg4_scatter python tools/agg_one.py cfg4 100000000 > $O/ncu_cfg4s.log 2>&1
$N -k kq_agg_partition_reduce -s 3 -o $O/agg_cfg4_reduce python tools/agg_one.py cfg4 100000000 > $O/ncu_cfg4r.log 2>&1
$N -k kq_filter_project -s 3 -o $O/filter_project python tools/fp_one.py 100000000 full > $O/ncu_fp.log 2>&1
$N -k regex:k_csv_fields -s 3 -o $O/csv_fields python tools/csv_one.py 10000000 > $O/ncu_csv.log 2>&1
ls -la $O | head -40
