#!/bin/bash
# Last GPU call of the round: parity, smoke, the csv bench line (device-resident + reader e2e) and its launch list.
O=gpurun_out/csvlast; mkdir -p $O
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 40 python -m pytest tests/test_gpu_csv.py -x -q -m gpu > $O/t_default.log 2>&1; el "tests: $(tail -1 $O/t_default.log)"
timeout 25 python __graft_entry__.py smoke > $O/smoke.log 2>&1; el "smoke: $(tail -1 $O/smoke.log | cut -c1-160)"
timeout 40 python bench.py --workload csv --steps 5 --warmup 3 > $O/bench_csv.json 2> $O/bench_csv.err; el "bench csv: $(python -c "import json; d=json.load(open('$O/bench_csv.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'])" 2>&1 | tail -1)"
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_csv.csv python bench.py --workload csv --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches_csv.log 2>&1; el "ncu launch list: $(wc -l < $O/launches_csv.csv) lines"
