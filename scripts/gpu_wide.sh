#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -12
run() { timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline_closest']; rs=d['roofline_shadow']; print(round(d['value'],1), 'build', round(d['bvh_build_ms'],2), {k:round(v,1) for k,v in d['stage_ms'].items()}, 'nodes/ray', round(r['nodes_per_ray'],1), 'prims/ray', round(r['prims_per_ray'],1), 'shadow nodes', round(rs['nodes_per_ray'],1))
    elif 'rror' in l: print(l.strip())"; }
echo wide; run
echo binary; MOX_FORCE_BINARY=1 run
