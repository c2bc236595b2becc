#!/bin/bash
# Round-2 evidence for profiles/: launch list of the bench at its own batch size (4 spp per wavefront) and
# `ncu --set full` captures (+ lts__t_bytes) of the hot kernels of one step.  Numbers printed by bench.py under ncu
# are not bench values.  Only CSV pages travel back (gpurun_out/ is limited to 64 MiB); a report is kept when small.
mkdir -p gpurun_out
TAG=${TAG:-r2}
if [ "$LIST" != "0" ]; then
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1
grep -c k_traverse gpurun_out/launches_$TAG.csv
fi
M=lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum
# (a) the two traversal kernels, second step of the first context (a step has 11 traversal launches), with source
timeout 1800 ncu --set full --metrics $M --clock-control none --import-source on -k regex:'k_traverse_wide' -s ${SKIP_T:-11} -c ${COUNT_T:-6} \
  -f -o /tmp/prof_trav python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_t_$TAG.log 2>&1
ncu -i /tmp/prof_trav.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_trav_raw.csv
ncu -i /tmp/prof_trav.ncu-rep --page source --csv --kernel-name regex:'k_traverse_wide' --launch-skip 0 --launch-count 2 > gpurun_out/prof_${TAG}_trav_source.csv 2>/dev/null
# (b) the other kernels of a step (18 launches: generate, classify / shade / apply per depth, accumulate)
timeout 1800 ncu --set full --metrics $M --clock-control none -k regex:'k_shade_disney|k_classify|k_apply|k_accumulate|k_generate' -s ${SKIP_S:-18} -c ${COUNT_S:-8} \
  -f -o /tmp/prof_shade python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_s_$TAG.log 2>&1
ncu -i /tmp/prof_shade.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_shade_raw.csv
ls -la /tmp/prof_trav.ncu-rep /tmp/prof_shade.ncu-rep
for f in trav shade; do s=$(stat -c %s /tmp/prof_$f.ncu-rep); if [ "$s" -lt 25000000 ]; then cp /tmp/prof_$f.ncu-rep gpurun_out/prof_${TAG}_$f.ncu-rep; fi; done
du -sh gpurun_out
