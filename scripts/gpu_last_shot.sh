#!/bin/bash
# One GPU call: bench the in-tree library and the candidates in variants/, then run the whole -m gpu suite on the best
# candidate (MOX_GPU_LIB) if it beats the in-tree library by more than 0.2 %.
mkdir -p gpurun_out
run() { timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ls_$1.json 2> gpurun_out/ls_$1.err; python scripts/show_bench.py gpurun_out/ls_$1.json $1; }
unset MOX_GPU_LIB; run def
for f in variants/*.so; do n=$(basename $f .so); MOX_GPU_LIB=$PWD/$f run $n; done
best=$(python - <<'P'
import json, glob, os
v = {}
for f in glob.glob("gpurun_out/ls_*.json"):
    try: v[os.path.basename(f)[3:-5]] = json.loads(open(f).read().strip().splitlines()[-1])["value"]
    except Exception: pass
d = v.pop("def", 0)
b = max(v, key=v.get) if v else ""
print(b if b and v[b] > 1.002 * d else "")
P
)
echo "best candidate: '$best'"
if [ -n "$best" ]; then
  MOX_GPU_LIB=$PWD/variants/$best.so timeout 300 python -m pytest tests -m gpu -q --timeout 280 -x 2>&1 | tail -4 | tee gpurun_out/ls_tests_$best.log
fi
