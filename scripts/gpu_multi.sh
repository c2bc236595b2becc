#!/bin/bash
N=${1:-2}
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 2>&1 | grep -E "^\{|rror|Traceback" | cut -c1-1600
