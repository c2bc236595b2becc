"""Worker for test_peer_memory_gather_across_processes: one rank of a 2-GPU NCCL job.  Each rank renders its tiles
on its own GPU, TileGather moves them into rank 0's gather buffer over peer memory, rank 0 compares the frame with
a 1-rank render."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import minimaloptix_b200 as mox  # noqa: E402
from minimaloptix_b200 import host  # noqa: E402
from minimaloptix_b200.parallel import TileGather  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    w, h, spp, seed = 300, 170, 3, 99     # ragged tiles
    sc = host.Scene.builtin("interior", 20000)
    api = host.ApiTable(mox.GPU_LIB, "mox_")
    ctx = mox.gpu().context(rank)
    sc.upload(api, ctx, w, h, 5)
    ctx.set_partition(rank, world, 32)
    ctx.build_accel()
    g = TileGather(ctx, rank, world, dev)
    ok = True
    frames = []
    for k in range(3):                    # three frames: both gather buffers get used, the schedule continues
        ctx.render(spp, seed)
        g.gather()
        if rank == 0:
            frames.append(g.read().copy())
    if rank == 0:
        full = mox.gpu().context(0)
        sc.upload(api, full, w, h, 5)
        full.build_accel()
        for k in range(3):
            full.render(spp, seed)
            ok = ok and np.array_equal(frames[k].view(np.uint32), full.read_accum().view(np.uint32))
        print("gathered frames bit-identical:", ok, "transport:", g.transport())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
