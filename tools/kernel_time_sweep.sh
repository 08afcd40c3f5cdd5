run() { env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k kq_hash_aggregate -s 3 -c 1 python tools/agg_one.py cfg3 100000000 2>&1 | grep -E "time_duration" | tr "\n" " "; echo " $*"; }
run KQ_CONSUMER_WAIT=1
run KQ_AGG_GEOM=12,6 KQ_AGG_RING_KB=100
run KQ_AGG_GEOM=14,6 KQ_AGG_RING_KB=110
run KQ_AGG_GEOM=16,6 KQ_AGG_RING_KB=120
run KQ_AGG_GEOM=12,5 KQ_AGG_RING_KB=90
run KQ_AGG_GEOM=16,5 KQ_AGG_RING_KB=110
run KQ_AGG_GEOM=16,4 KQ_AGG_RING_KB=100
run KQ_AGG_GEOM=12,7 KQ_AGG_RING_KB=80
run KQ_AGG_GEOM=10,6 KQ_AGG_RING_KB=100
run KQ_AGG_GEOM=8,6 KQ_AGG_RING_KB=110
