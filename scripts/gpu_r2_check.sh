#!/bin/bash
# GPU check of the working tree: the -m gpu suite, then bench.py (full JSON line kept in gpurun_out/bench_$TAG.json)
mkdir -p gpurun_out
TAG=${TAG:-check}
if [ "$RUN_TESTS" != "0" ]; then
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -25 > gpurun_out/tests_$TAG.log; tail -25 gpurun_out/tests_$TAG.log
fi
timeout 900 python bench.py --steps ${STEPS:-8} --warmup 3 $BENCH_ARGS > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || tail -20 gpurun_out/bench_$TAG.err
python scripts/show_bench.py gpurun_out/bench_$TAG.json $TAG
python - <<P
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
for k in ("roofline","roofline_closest","roofline_build","serial_pass"):
    v=d.get(k) or {}
    print(k,{a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ("kernel","achieved","frac","bytes_per_ray","fixed_bytes_per_ray","nodes_per_ray","prims_per_ray","avg_launch_ms","avg_launch_ms_in_timed_region","share_of_step","ms","ms_per_step","bounce_only_extend_mrays_per_s","primary_extend_mrays_per_s","stage_ms_per_step")})
for r in d.get("per_depth") or []: print(r)
print("build_ms",d["bvh_build_ms"],"cpu",d.get("cpu_baseline"))
P
