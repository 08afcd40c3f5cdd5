mkdir -p gpurun_out/r2
bash tools/ncu_agg.sh cfg3 1000000 fe1m
