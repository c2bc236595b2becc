#!/bin/bash
# Final evidence for profiles/: launch list of one bench step + full captures of the hot kernels.
mkdir -p gpurun_out
export MOX_MAX_BATCH_PATHS=8400000   # 1 spp per wavefront keeps the ncu replays short
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_traverse|k_shade_disney|k_logic|k_apply' -s 6 -c 10 -f -o gpurun_out/prof_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log | cut -c1-200
ls -la gpurun_out | tail -5
