#!/bin/bash
# tools/fe_sweep.sh <workload> <geom>... : device-resident rows/s of the low-cardinality aggregate kernel per tile geometry (KQ_AGG_GEOM="rows,warps[,stages]")
wl=$1; shift
for g in "$@"; do
  if [ "$g" = "default" ]; then unset KQ_AGG_GEOM; else export KQ_AGG_GEOM=$g; fi
  out=$(timeout 180 python bench.py --workload $wl --steps 10 --no-sub --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$wl $g $(echo "$out" | python -c 'import sys,json
try:
    l=json.loads(sys.stdin.read()); print("rows/s %.4g  ms %.3f  frac %.3f  check %s" % (l["value"], l["ms_per_step"], l["roofline"]["frac"], l["check"].get("count_ok")))
except Exception as e: print("FAILED", e)')"
done
