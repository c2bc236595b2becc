#!/bin/bash
# All five BASELINE configs on one B200 next to the oracle (gpurun_out/configs.json), and the ncu capture of the
# incoherent-ray traversal on the 10 M-triangle soup (config 5: the one case where the scene exceeds L2).
mkdir -p gpurun_out
timeout 1500 python scripts/run_configs.py > gpurun_out/run_configs.log 2>&1; tail -3 gpurun_out/run_configs.log | cut -c1-600
timeout 900 ncu --set full --metrics lts__t_bytes.sum --clock-control none -k regex:'k_traverse_wide' -s 1 -c 1 -f -o /tmp/prof_soup \
  python scripts/soup_trace.py 10000000 24 > gpurun_out/ncu_soup.log 2>&1
ncu -i /tmp/prof_soup.ncu-rep --page raw --csv > gpurun_out/prof_r2_soup_raw.csv
tail -3 gpurun_out/ncu_soup.log
rm -f gpurun_out/c?_*.png
