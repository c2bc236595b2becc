#!/bin/bash
# A/B: bench the default build against the libraries in variants/ (and env switches)
run() { timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), 'build', round(d['bvh_build_ms'],2), {k:round(v,1) for k,v in d['stage_ms'].items()})
    elif 'rror' in l: print(l.strip())"; }
echo default; run
for e in $AB_ENVS; do echo "env $e"; env $e bash -c "$(declare -f run); run"; done
for f in variants/*.so; do [ -e "$f" ] || continue; echo $f; MOX_GPU_LIB=$PWD/$f run; done
