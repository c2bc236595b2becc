#!/bin/bash
# round 2, first GPU call: the whole GPU suite incl. the new full-resolution parity cases, then a short bench
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -rA 2>&1 | grep -E "passed|failed|PASSED|FAILED|rmse|Error|error|assert" | tail -80 > gpurun_out/r2_tests.log; tail -60 gpurun_out/r2_tests.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
