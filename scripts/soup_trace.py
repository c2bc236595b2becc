#!/usr/bin/env python3
"""ncu target for BASELINE config 5: build the N-triangle soup, trace 2^LOGN incoherent rays twice (the second launch
is the one to capture: -k regex:k_traverse_wide -s 1 -c 1).  usage: soup_trace.py [N] [LOGN]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import minimaloptix_b200 as mox
from minimaloptix_b200 import host
n_tris = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
sc = host.Scene.builtin("soup", n_tris)
g = mox.gpu().context(0)
sc.upload(host.ApiTable(mox.GPU_LIB, "mox_"), g, 64, 64, 5)
print("build ms", g.build_accel())
gen = torch.Generator(device="cuda").manual_seed(12345)
r = torch.empty((n, 8), device="cuda")
r[:, 0:3] = torch.rand((n, 3), generator=gen, device="cuda")
d = torch.randn((n, 3), generator=gen, device="cuda")
r[:, 4:7] = d / d.norm(dim=1, keepdim=True)
r[:, 3], r[:, 7] = 1e-3, 1e27
hits = torch.empty((n, 4), device="cuda")
torch.cuda.synchronize()
for _ in range(2):
    print("trace ms", g.trace_closest_device(r.data_ptr(), n, hits.data_ptr()))
