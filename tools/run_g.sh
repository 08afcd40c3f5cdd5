echo "== big buckets (no spill)"; KQ_PART_CAP=12000 timeout 300 python tools/part_loop.py 4 2>&1 | tail -4
echo "== grid 1"; KQ_PART_GRID=1 timeout 300 python tools/part_loop.py 3 2>&1 | tail -6
echo "== grid 2"; KQ_PART_GRID=2 timeout 300 python tools/part_loop.py 3 2>&1 | tail -6
echo "== hot 0.0"; timeout 300 python tools/part_loop.py 3 0.0 2>&1 | tail -4
echo "== hot 0.02"; timeout 300 python tools/part_loop.py 3 0.02 2>&1 | tail -6
