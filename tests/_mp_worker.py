"""Worker for tests/test_multiprocess.py: one rank of a world_size-N gloo job on CPU.  Each rank
renders its tiles with a CPU context, the tiles are gathered to rank 0 through
minimaloptix_b200.parallel.TileGather, and rank 0 compares against a 1-rank render."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402  (tests may use the oracle)
from minimaloptix_b200 import host  # noqa: E402
from minimaloptix_b200.parallel import TileGather  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, spp, tile, seed = 100, 70, 2, 16, 4242  # ragged tiles on both axes
    sc = host.Scene.builtin("spheres_lens")
    api = host.ApiTable(oracle.ORACLE_LIB, "orc_")
    ctx = oracle.context(threads=2)
    sc.upload(api, ctx, w, h, 4)
    ctx.set_partition(rank, world, tile)
    ctx.build_accel()
    ctx.render(spp, seed)
    mine = ctx.read_accum()
    own = (mine != 0).any(axis=2).sum()
    g = TileGather(ctx, rank, world, torch.device("cpu"))
    assert own <= g.owned[rank]
    assert sum(g.owned) == w * h
    g.gather()
    ok = True
    if rank == 0:
        full = oracle.context(threads=2)
        sc.upload(api, full, w, h, 4)
        full.build_accel()
        full.render(spp, seed)
        want = full.read_accum()
        got = ctx.read_accum()
        ok = np.array_equal(got.view(np.uint32), want.view(np.uint32))
        print("gathered image bit-identical:", ok, "wire bytes", g.bytes_on_the_wire())
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
