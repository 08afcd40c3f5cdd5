mkdir -p gpurun_out/r2
echo "== default"; timeout 300 python tools/part_loop.py 12 2>&1 | tail -12
echo "== scalar merge"; KQ_PART_SCALAR_MERGE=1 timeout 300 python tools/part_loop.py 12 2>&1 | tail -8
echo "== no partition"; KQ_NO_PARTITION=1 timeout 300 python tools/part_loop.py 6 2>&1 | tail -8
echo "== hot 0.5"; timeout 300 python tools/part_loop.py 8 0.5 2>&1 | tail -8
