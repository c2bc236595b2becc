#!/bin/bash
# quick GPU check: tests + short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x 2>&1 | tail -15
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
