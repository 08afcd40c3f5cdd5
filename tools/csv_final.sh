#!/bin/bash
# Second (last) GPU call for the CSV work of this round: parity of the final kernels, smoke, the batch-size A/B of the
# deferred quote-block work, the launch list and the bench line of the csv workload, and the default bench line as a check
# that nothing else moved. Ordered by importance; every step has its own timeout.
O=gpurun_out/csvfinal; mkdir -p $O
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 60 python -m pytest tests/test_gpu_csv.py -x -q -m gpu > $O/t_default.log 2>&1; el "tests: $(tail -1 $O/t_default.log)"
timeout 30 python __graft_entry__.py smoke > $O/smoke.log 2>&1; el "smoke: $(tail -1 $O/smoke.log | cut -c1-160)"
B="python bench.py --workload csv --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
for b in 16 1 32 8; do KQ_CSV_BATCH=$b timeout 30 $B > $O/b_batch$b.json 2> $O/b_batch$b.err; el "bench batch $b: $(python -c "import json; d=json.load(open('$O/b_batch$b.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'])" 2>&1 | tail -1)"; done
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_csv.csv python bench.py --workload csv --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches_csv.log 2>&1; el "ncu launch list: $(wc -l < $O/launches_csv.csv) lines"
timeout 30 python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline > $O/b_default.json 2> $O/b_default.err; el "default bench: $(python -c "import json; d=json.load(open('$O/b_default.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['check'])" 2>&1 | tail -1 | cut -c1-200)"
timeout 70 python bench.py --workload csv --steps 5 --warmup 3 > $O/bench_csv.json 2> $O/bench_csv.err; el "bench csv full: $(cut -c1-120 $O/bench_csv.json)"
