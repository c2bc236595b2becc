#!/bin/bash
# A/B of libraries and environment switches in one GPU call.  CASES = "name|lib|ENV=1,ENV2=x ..." (lib "default" = in-tree)
mkdir -p gpurun_out
if [ "$RUN_TESTS" = "1" ]; then timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -25 > gpurun_out/ab_tests.log; tail -5 gpurun_out/ab_tests.log; fi
for c in $CASES; do
  name=${c%%|*}; rest=${c#*|}; lib=${rest%%|*}; envs=${rest#*|}
  ( [ "$lib" = "default" ] || export MOX_GPU_LIB=$PWD/$lib
    for e in ${envs//,/ }; do [ -n "$e" ] && export "$e"; done
    timeout 600 python bench.py --steps ${STEPS:-6} --warmup 3 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || tail -5 gpurun_out/ab_$name.err )
  python scripts/show_bench.py gpurun_out/ab_$name.json $name
done
