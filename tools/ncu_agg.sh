#!/bin/bash
# tools/ncu_agg.sh <workload> <rows> <tag>: one ncu --set full capture of the aggregate kernel (after a plain run exited 0) + the launch list
wl=$1; rows=$2; tag=$3
python bench.py --workload $wl --rows $rows --steps 2 --no-sub --no-e2e --no-cpu-baseline > gpurun_out/r2/plain_$tag.json 2> gpurun_out/r2/plain_$tag.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:kq_group_aggregate -c 1 -s 3 -o gpurun_out/r2/ncu_$tag -f python bench.py --workload $wl --rows $rows --steps 2 --no-sub --no-e2e --no-cpu-baseline > gpurun_out/r2/ncu_$tag.log 2>&1
ncu -i gpurun_out/r2/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/r2/ncu_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2/ncu_$tag.ncu-rep --page source --csv > gpurun_out/r2/ncu_${tag}_src.csv 2>/dev/null
ls -la gpurun_out/r2/ | grep $tag
