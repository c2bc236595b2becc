"""Multi-GPU plumbing for one process per GPU (torch.distributed): the image is tile-split with
mox_set_partition, the scene is replicated, and the only exchange is the gather of each rank's owned
pixels on rank 0 (SURVEY.md §8 e).  The reference is single-GPU (one optix::Context,
MinimalOptiX.cpp:131), so there is nothing to mirror here.

Two transports:
  * peer memory (default on GPUs): rank 0 exports its two gather buffers as CUDA IPC handles once;
    per frame every rank writes the pixels it owns straight into rank 0's buffer with NVLink stores
    (mox_gather_push) and a barrier closes the frame — no staging buffer, no collective on the data
    path, no unpack kernels.  Rank 0 then reads the frame asynchronously (read_begin / read_end), two
    buffers alternating so the device->host copy of frame k overlaps the rendering of frame k+1.
  * collective (gloo on CPU for the tests, or MOX_GATHER=nccl): pack -> dist.gather -> unpack into
    rank 0's accumulation buffer.

The in-process alternative — one handle, several GPUs, host threads in C++ — is mox_create_multi
(include/mox.h); it shares the push kernel and the gather buffers with the peer-memory transport."""
import os

import torch
import torch.distributed as dist

from ._binding import MoxError


class TileGather:
    """Gathers the accumulation tiles of every rank into one image on rank 0."""

    def __init__(self, ctx, rank, world, device):
        self.ctx, self.rank, self.world, self.device = ctx, rank, world, device
        self.owned = [ctx.owned_pixels(r) for r in range(world)]
        self.which, self.last = 0, None
        self.p2p = False
        gpu = torch.device(device).type == "cuda" and hasattr(ctx.b, "gather_export")
        if world > 1 and gpu and os.environ.get("MOX_GATHER", "p2p") != "nccl":
            self.p2p = self._setup_peer_memory()
        self.pack = self.recv = None
        if not self.p2p:
            pad = max(self.owned) if self.owned else 0
            self.pack = torch.zeros(max(pad, 1) * 3, dtype=torch.float32, device=device)
            if world > 1 and rank == 0:
                self.recv = [torch.zeros_like(self.pack) for _ in range(world)]

    def _setup_peer_memory(self):
        """Collective.  True when every rank mapped rank 0's gather buffers."""
        handles = [None, None]
        ok = 1
        try:
            if self.rank == 0:
                handles = [self.ctx.gather_export(0), self.ctx.gather_export(1)]
        except MoxError:
            ok = 0
        dist.broadcast_object_list(handles, src=0)
        if self.rank != 0:
            try:
                for w in (0, 1):
                    if handles[w] is None:
                        raise MoxError("rank 0 could not export its gather buffer")
                    self.ctx.gather_import(w, handles[w])
            except MoxError:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    def transport(self):
        return "peer-memory stores (CUDA IPC over NVLink)" if self.p2p else ("none" if self.world == 1 else "collective gather")

    def bytes_on_the_wire(self):
        """Payload rank 0 receives: 12 bytes per pixel owned by the other ranks."""
        return 12 * sum(self.owned[1:])

    def gather(self):
        """Collective: call on every rank.  Afterwards the frame is complete on rank 0 — in its gather
        buffer (peer memory; read it with read_begin/read_end) or in its accumulation buffer."""
        if self.world == 1:
            return
        if self.p2p:
            self.ctx.gather_push(self.which)     # synchronous: this rank's pixels have landed
            dist.barrier()                       # ... and so have everyone else's
            self.last, self.which = self.which, self.which ^ 1
            return
        self.ctx.pack_owned(self.pack.data_ptr())
        dist.gather(self.pack, self.recv, dst=0)
        if self.rank == 0:
            if self.pack.is_cuda:
                torch.cuda.synchronize(self.device)
            for r in range(1, self.world):
                self.ctx.unpack_owned(r, self.recv[r].data_ptr())

    # ---- rank 0: the gathered frame on the host
    def read_begin(self):
        """Start the device->host copy of the frame gathered last (returns at once on GPUs)."""
        if self.p2p:
            self._pending = self.last
            self.ctx.read_gathered_begin(self.last)
        else:
            self.ctx.read_accum_begin()

    def read_end(self):
        """(H, W, 3) view of the pinned host image whose copy read_begin started."""
        if self.p2p:
            return self.ctx.read_gathered_end(self._pending)
        return self.ctx.read_accum_end()

    def read(self):
        self.read_begin()
        return self.read_end()
