#!/bin/bash
# Kernel-only time of kq_hash_aggregate (ncu gpu__time_duration, 100 M rows of config 3) for tuning variables given as
# env assignments, one variant per line of the here-doc below. Used for the experiments recorded in DESIGN.md §5.
run() { env "$@" ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k kq_hash_aggregate -s 3 -c 1 python tools/agg_one.py cfg3 100000000 2>&1 | grep -E "time_duration|inst_executed" | tr "\n" " "; echo " $*"; }
run KQ_AGG_GEOM=4,7 KQ_AGG_RING_KB=32
run KQ_AGG_GEOM=4,7 KQ_AGG_RING_KB=45
run KQ_AGG_GEOM=4,7 KQ_AGG_RING_KB=64
