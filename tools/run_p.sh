mkdir -p gpurun_out/r2
bash tools/fe_sweep.sh cfg3 6,7,3 6,7 4,7,4 8,7,2,1 10,7,2 | tee gpurun_out/r2/sweep3n.log
bash tools/fe_sweep.sh cfg5 8,7,2 6,7,3 6,6,3 | tee gpurun_out/r2/sweep5n.log
