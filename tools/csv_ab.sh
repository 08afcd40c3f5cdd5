#!/bin/bash
# One GPU call for the CSV work of this round: parity of kq_csv_scan / kq_csv_reader_* on the device (both kernel variants),
# smoke, and the A/B of the variants on the csv workload (10 M records, device-resident text). Ordered by importance; every
# step has its own timeout so that the call ends inside the GPU budget that is left.
O=gpurun_out/csvab; mkdir -p $O
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 80 python -m pytest tests/test_gpu_csv.py -x -q -m gpu > $O/t_default.log 2>&1; el "tests default: $(tail -1 $O/t_default.log)"
timeout 40 python __graft_entry__.py smoke > $O/smoke.log 2>&1; el "smoke: $(tail -1 $O/smoke.log | cut -c1-160)"
B="python bench.py --workload csv --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { KQ_CSV_FIELDS=$1 KQ_CSV_MASKS=$2 timeout 40 $B > $O/b_$1_$2.json 2> $O/b_$1_$2.err; el "bench $1 $2: $(python -c "import json,sys; d=json.load(open('$O/b_$1_$2.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'])" 2>&1 | tail -1)"; }
run record recompute
run field stored
KQ_CSV_FIELDS=field KQ_CSV_MASKS=stored timeout 60 python -m pytest tests/test_gpu_csv.py -x -q -m gpu > $O/t_variant.log 2>&1; el "tests field+stored: $(tail -1 $O/t_variant.log)"
run field recompute
run record stored
timeout 50 python bench.py --workload csv --steps 5 --warmup 3 --no-cpu-baseline > $O/b_e2e.json 2> $O/b_e2e.err; el "bench e2e (reader): $(python -c "import json; d=json.load(open('$O/b_e2e.json')); print(d['value'], d['e2e'])" 2>&1 | tail -1 | cut -c1-300)"
