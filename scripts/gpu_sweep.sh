#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x 2>&1 | tail -5
run() { timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(round(d['value'],1), 'build_ms', round(d['bvh_build_ms'],2), {k:round(v,1) for k,v in d['stage_ms'].items()}, 'nodes/ray', round(r['nodes_per_ray'],1), 'prims/ray', round(r['prims_per_ray'],1))
    elif 'rror' in l: print(l.strip())"; }
echo LBVH; MOX_FORCE_LBVH=1 run
for r in 4 8 16 32; do echo "PLOC r=$r"; MOX_PLOC_RADIUS=$r run; done
