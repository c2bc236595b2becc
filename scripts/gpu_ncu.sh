#!/bin/bash
mkdir -p gpurun_out
export MOX_MAX_BATCH_PATHS=8400000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_traverse' -s 4 -c 4 -f -o gpurun_out/prof_trav python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
