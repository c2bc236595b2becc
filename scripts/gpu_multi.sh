#!/bin/bash
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-6} --warmup 3 2>&1 | grep -E "^\{|rror|Traceback" | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'wall/step', round(d['wall_ms_per_step'],2), 'gather_ms', round(d['gather_ms'],2), 'render_ms_rank0', round(d['render_ms_rank0'],1), 'per_rank', d.get('render_ms_per_rank'), 'tile', d.get('tile'), {k:round(v,1) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['value'],1))
    else: print(l.strip())"
