#!/bin/bash
# Multi-GPU evidence on one 8-GPU box: the tests that need >= 2 GPUs, bench.py under torchrun at N = 2, 4, 8
# (peer-memory gather), and BASELINE config 4 as it is worded — 1024 spp at 4K on 8 GPUs through the C++ path
# (mox_cli --gpus 8, power-of-two snapshots).
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "multi_handle or peer_memory or asynchronous" 2>&1 | tail -3
timeout 600 python bench.py --steps ${STEPS:-8} --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
python scripts/show_bench.py gpurun_out/scale_1.json N1
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps ${STEPS:-8} --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err || tail -20 gpurun_out/scale_$N.err
  python scripts/show_bench.py gpurun_out/scale_$N.json N$N
done
timeout 900 ./minimaloptix_b200/mox_cli --scene interior --width 3840 --height 2160 --max-depth 5 --spp 1024 --seed 13738406 --gpus 8 --snapshots --out /tmp/c4_8gpu | tail -1 | tee gpurun_out/cli_c4_1024spp_8gpu.json
timeout 900 ./minimaloptix_b200/mox_cli --scene interior --width 3840 --height 2160 --max-depth 5 --spp 1024 --seed 13738406 --gpus 8 --out /tmp/c4_8gpu_nosnap | tail -1 | tee gpurun_out/cli_c4_1024spp_8gpu_nosnap.json
cmp /tmp/c4_8gpu.png /tmp/c4_8gpu_nosnap.png && echo "final images identical with and without snapshots"
python - <<'P'
import zlib,sys
# keep a small copy of the converged frame (downscaled 4x) as evidence
try:
    sys.path.insert(0,'.')
    from minimaloptix_b200 import host
except Exception as e:
    print(e)
P
ls -la /tmp/c4_8gpu*.png | head -14
