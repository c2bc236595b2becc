#!/bin/bash
# Round-2 evidence for profiles/: launch list of the bench at its own batch size (4 spp per wavefront) and one
# `ncu --set full` capture (+ lts__t_bytes) of the hot kernels of one step.  Numbers printed by bench.py under ncu are
# not bench values.
mkdir -p gpurun_out
TAG=${TAG:-r2}
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1
grep -c k_traverse gpurun_out/launches_$TAG.csv
# second step of the first context: skip the kernels of the build + the warm-up step (counted from the launch list)
SKIP=${SKIP:-40}
timeout 2400 ncu --set full --metrics lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum --clock-control none --import-source on \
  -k regex:'k_traverse_wide|k_shade_disney|k_classify|k_apply|k_accumulate|k_generate' -s $SKIP -c ${COUNT:-28} -f -o gpurun_out/prof_$TAG \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_f_$TAG.log 2>&1
tail -2 gpurun_out/ncu_f_$TAG.log | cut -c1-200
ls -la gpurun_out/prof_$TAG.ncu-rep
