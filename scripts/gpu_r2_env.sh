#!/bin/bash
# A/B by environment: each item of VARIANTS is "name:ENV1=a,ENV2=b" (or "name:"), run through bench.py
mkdir -p gpurun_out
if [ "$RUN_TESTS" = "1" ]; then
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -25 > gpurun_out/ab_tests.log; tail -25 gpurun_out/ab_tests.log
fi
for v in $VARIANTS; do
  name=${v%%:*}; envs=${v#*:}
  ( IFS=','; for e in $envs; do export "$e"; done; unset IFS
    timeout 600 python bench.py --steps ${STEPS:-6} --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/env_$name.json 2> gpurun_out/env_$name.err || tail -5 gpurun_out/env_$name.err )
  python scripts/show_bench.py gpurun_out/env_$name.json $name
done
