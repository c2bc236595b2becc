#!/bin/bash
run() { timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), {k:round(v,1) for k,v in d['stage_ms'].items()})
    elif 'rror' in l: print(l.strip())"; }
echo nosort; run
echo sort; MOX_SORT_RAYS=1 run
MOX_SORT_RAYS=1 timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "render or batched or partition" 2>&1 | tail -3
