#!/bin/bash
# A/B on the random-spheres config (analytic primitives only) and the bench scene
echo "C2 default"; python scripts/run_configs.py --only C2 2>&1 | tail -1
for f in variants/*.so; do echo "C2 $f"; MOX_GPU_LIB=$PWD/$f python scripts/run_configs.py --only C2 2>&1 | tail -1; done
bash scripts/gpu_ab.sh
