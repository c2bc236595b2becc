#!/bin/bash
run() { timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), {k:round(v,1) for k,v in d['stage_ms'].items()})
    elif 'rror' in l: print(l.strip())"; }
python -c "
import torch
p=torch.cuda.get_device_properties(0); print('L2', p.L2_cache_size)
import ctypes
" 
echo persist; run
echo nopersist; MOX_L2_PERSIST=0 run
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -2
