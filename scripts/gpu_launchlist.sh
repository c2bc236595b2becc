#!/bin/bash
mkdir -p gpurun_out
export MOX_MAX_BATCH_PATHS=8400000
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_new.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_l.log | cut -c1-100
