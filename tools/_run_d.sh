timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in c2 c1; do timeout 120 python tools/stage_time.py $c 12 2>&1 | tail -1; done
