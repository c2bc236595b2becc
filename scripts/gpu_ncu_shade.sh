#!/bin/bash
mkdir -p gpurun_out
export MOX_MAX_BATCH_PATHS=8400000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_shade_disney|k_logic|k_apply|k_generate|k_accumulate' -s 1 -c 5 -f -o gpurun_out/prof_shade python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_shade.log 2>&1
tail -2 gpurun_out/ncu_shade.log | cut -c1-200
