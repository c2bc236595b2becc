#!/bin/bash
# multi-GPU checks: the tests that need >= 2 GPUs, then bench.py under torchrun at N = $NGPU (peer-memory gather and,
# for comparison, the collective gather)
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "multi_handle or peer_memory or asynchronous" 2>&1 | tail -15
for mode in p2p nccl; do
  MOX_GATHER=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-8} --warmup 3 --no-cpu-baseline > gpurun_out/multi_${N}_$mode.json 2> gpurun_out/multi_${N}_$mode.err || tail -20 gpurun_out/multi_${N}_$mode.err
  python scripts/show_bench.py gpurun_out/multi_${N}_$mode.json N${N}_$mode
done
# the C++ path: one process, N devices, no Python
timeout 600 ./minimaloptix_b200/mox_cli --scene interior --width 3840 --height 2160 --max-depth 5 --spp 32 --seed 13738406 --gpus $N --out gpurun_out/cli_interior_$N | tail -1
timeout 600 ./minimaloptix_b200/mox_cli --scene interior --width 3840 --height 2160 --max-depth 5 --spp 32 --seed 13738406 --device 0 --out gpurun_out/cli_interior_1 | tail -1
cmp gpurun_out/cli_interior_$N.png gpurun_out/cli_interior_1.png && echo "cli images identical"
rm -f gpurun_out/cli_interior_*.png
