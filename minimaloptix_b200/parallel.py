"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink), the image is
tile-split with mox_set_partition, the scene is replicated, and the only exchange is the final
gather of each rank's owned pixels to rank 0 (SURVEY.md §8 e).  The reference is single-GPU
(one optix::Context, MinimalOptiX.cpp:131), so there is nothing to mirror here.

The same code runs on CPU tensors with the gloo backend (used by the world_size-2 tests with a
CPU context whose pack/unpack take host pointers)."""
import torch
import torch.distributed as dist


class TileGather:
    """Reusable buffers for gathering the accumulation tiles of every rank into rank 0's image."""

    def __init__(self, ctx, rank, world, device):
        self.ctx, self.rank, self.world, self.device = ctx, rank, world, device
        self.owned = [ctx.owned_pixels(r) for r in range(world)]
        pad = max(self.owned) if self.owned else 0
        self.pack = torch.zeros(max(pad, 1) * 3, dtype=torch.float32, device=device)
        self.recv = None
        if world > 1 and rank == 0:
            self.recv = [torch.zeros_like(self.pack) for _ in range(world)]

    def bytes_on_the_wire(self):
        """Payload rank 0 receives: 12 bytes per pixel owned by the other ranks."""
        return 12 * sum(self.owned[1:])

    def gather(self):
        """Collective: call on every rank.  After it rank 0's accumulation buffer holds all tiles."""
        if self.world == 1:
            return
        self.ctx.pack_owned(self.pack.data_ptr())
        dist.gather(self.pack, self.recv, dst=0)
        if self.rank == 0:
            if self.pack.is_cuda:
                torch.cuda.synchronize(self.device)
            for r in range(1, self.world):
                self.ctx.unpack_owned(r, self.recv[r].data_ptr())
