echo "== with drain"; timeout 300 python tools/part_loop.py 12 2>&1 | grep "bad runs\|^run" | tail -5
echo "== without drain"; KQ_NO_STAGE_DRAIN=1 timeout 300 python tools/part_loop.py 4 2>&1 | grep "bad runs\|^run" | tail -5
echo "== hot 0.5 with drain"; timeout 300 python tools/part_loop.py 6 0.5 2>&1 | grep "bad runs\|^run" | tail -5
