#!/usr/bin/env python3
"""Host scheduling noise on a GPU box: the largest gaps between iterations of a busy loop that does nothing else
(10 s), next to the same loop around a trivial CUDA call (cudaEventQuery through torch).  Used to tell host-side
stalls of bench.py's wall clock (steps of 150 ms next to 108 ms ones, device time unchanged) from anything the
library does."""
import time, sys
def gaps(fn, seconds):
    out = []
    t_end = time.perf_counter() + seconds
    last = time.perf_counter()
    while last < t_end:
        fn()
        now = time.perf_counter()
        if now - last > 2e-3:
            out.append(round((now - last) * 1e3, 1))
        last = now
    return out
print("pure host loop, gaps > 2 ms:", gaps(lambda: None, float(sys.argv[1]) if len(sys.argv) > 1 else 10.0))
try:
    import torch
    ev = torch.cuda.Event()
    ev.record()
    torch.cuda.synchronize()
    print("loop around cudaEventQuery, gaps > 2 ms:", gaps(ev.query, float(sys.argv[1]) if len(sys.argv) > 1 else 10.0))
except Exception as e:
    print("no cuda:", e)
