#!/usr/bin/env python3
"""Config 5 (10 M-triangle soup, incoherent rays): wide vs binary traversal, PLOC build."""
import os, sys, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import minimaloptix_b200 as mox
from minimaloptix_b200 import host, structs as S
n_tris = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
api = host.ApiTable(mox.GPU_LIB, "mox_")
sc = host.Scene.builtin("soup", n_tris)
res = {}
gen = torch.Generator(device="cuda").manual_seed(12345)
n = 1 << 24
r = torch.empty((n, 8), device="cuda")
r[:, 0:3] = torch.rand((n, 3), generator=gen, device="cuda")
d = torch.randn((n, 3), generator=gen, device="cuda")
r[:, 4:7] = d / d.norm(dim=1, keepdim=True)
r[:, 3], r[:, 7] = 1e-3, 1e27
hits = {}
for name, flags in (("wide", S.ACCEL_DEFAULT), ("binary", S.ACCEL_BINARY), ("lbvh", S.ACCEL_LBVH)):
    g = mox.gpu().context(0)
    sc.upload(api, g, 64, 64, 5)
    build = min(g.build_accel(flags) for _ in range(2))
    h = torch.empty((n, 4), device="cuda")
    torch.cuda.synchronize()
    ms = min(g.trace_closest_device(r.data_ptr(), n, h.data_ptr()) for _ in range(3))
    hits[name] = h.view(torch.int32).clone()
    st = g.stats()
    res[name] = {"build_ms": build, "ms": ms, "mrays_per_s": n / ms / 1e3, "nodes": st["n_nodes"], "node_bytes": st["node_bytes"]}
    del g
res["wide_vs_binary_rays_differing"] = int((hits["wide"] != hits["binary"]).any(dim=1).sum())
res["wide_vs_lbvh_rays_differing"] = int((hits["wide"] != hits["lbvh"]).any(dim=1).sum())
print(json.dumps(res))
