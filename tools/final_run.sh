#!/bin/bash
# Round-end measurement on one B200: tests, smoke, the bench lines, the launch lists and the ncu captures that profiles/ summarises.
O=gpurun_out/final; mkdir -p $O
python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; tail -2 $O/pytest.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err; cut -c1-300 $O/bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for w in cfg2i csv; do python bench.py --workload $w --steps 5 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; cut -c1-120 $O/bench_$w.json; done
python bench.py --workload cfg3 --rows 125000000 --no-sub --no-cpu-baseline > $O/bench_cfg3_125M.json 2> $O/bench_cfg3_125M.err; cut -c1-160 $O/bench_cfg3_125M.json
L="ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv"
$L --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-sub --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1
$L --log-file $O/launches_cfg5.csv python bench.py --workload cfg5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches5.log 2>&1
$L --log-file $O/launches_cfg2f.csv python bench.py --workload cfg2f --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches2f.log 2>&1
$L --log-file $O/launches_cfg4.csv python bench.py --workload cfg4 --rows 100000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches4.log 2>&1
$L --log-file $O/launches_csv.csv python bench.py --workload csv --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches_csv.log 2>&1
N="ncu --set full --clock-control none --import-source on -c 1"
B="python bench.py --steps 2 --no-sub --no-e2e --no-cpu-baseline"
$N -k regex:kq_group_aggregate -s 3 -o $O/agg_cfg3 -f $B --workload cfg3 > $O/ncu_cfg3.log 2>&1
$N -k regex:kq_group_aggregate -s 3 -o $O/agg_cfg5 -f $B --workload cfg5 > $O/ncu_cfg5.log 2>&1
$N -k regex:kq_filter_project -s 3 -o $O/filter_project -f $B --workload cfg2f > $O/ncu_fp.log 2>&1
$N -k regex:kq_hash_aggregate -s 3 -o $O/agg_cfg4_scatter -f $B --workload cfg4 --rows 100000000 > $O/ncu_cfg4s.log 2>&1
$N -k regex:kq_agg_partition_reduce -s 3 -o $O/agg_cfg4_reduce -f $B --workload cfg4 --rows 100000000 > $O/ncu_cfg4r.log 2>&1
for r in agg_cfg3 agg_cfg5 filter_project; do ncu -i $O/$r.ncu-rep --page source --csv > $O/${r}_src.csv 2>/dev/null; done
ls -la $O | head -50
