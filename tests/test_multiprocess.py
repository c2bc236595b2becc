"""world_size > 1 on CPU (gloo): the tile partition + gather plumbing used by bench.py --gpus N."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_tile_gather_over_gloo(world):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_mp_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} failed:\n{outs[r]}"
    assert "bit-identical: True" in outs[0]


def test_partition_covers_every_pixel_once(orc):
    """Pure host logic: for 1/2/4/8 ranks and ragged images every pixel has exactly one owner."""
    ctx = orc.context()
    for (w, h, tile) in [(64, 64, 32), (100, 70, 16), (1920, 1080, 32), (33, 17, 8)]:
        ctx.set_globals(w, h, 5)
        for world in (1, 2, 4, 8):
            total = 0
            for r in range(world):
                ctx.set_partition(r, world, tile)
                total += ctx.owned_pixels(r)
            assert total == w * h
            # balance: no rank owns more than 1.5x the mean + one tile (diagonal interleave)
            sizes = []
            for r in range(world):
                ctx.set_partition(r, world, tile)
                sizes.append(ctx.owned_pixels(r))
            assert max(sizes) <= 1.5 * (w * h / world) + tile * tile * 2
