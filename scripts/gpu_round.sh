#!/bin/bash
# Runs on the GPU box under gpurun: tests, bench, ncu launch list and one full capture of the traversal kernel.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -5
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 2 -c 2 -f -o gpurun_out/prof_extend python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
