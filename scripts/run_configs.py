#!/usr/bin/env python3
"""Run the five BASELINE.json configs on one B200 next to the CPU oracle (same run, host core
count printed) and write gpurun_out/configs.json.  Mrays/s counts extend rays (primary + bounce);
shadow rays are reported separately.  Oracle legs use a reduced, stated spp (Mrays/s is a rate)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import minimaloptix_b200 as mox  # noqa: E402
import oracle  # noqa: E402
from minimaloptix_b200 import host  # noqa: E402

GPU = host.ApiTable(mox.GPU_LIB, "mox_")
ORC = host.ApiTable(oracle.ORACLE_LIB, "orc_")
CORES = os.cpu_count()


def rays(st):
    return st["rays_primary"] + st["rays_bounce"]


def render_case(name, sc, w, h, spp, depth, seed, orc_spp, save=None):
    g = mox.gpu().context(0)
    sc.upload(GPU, g, w, h, depth)
    build = g.build_accel()
    g.render(1, seed ^ 0x55)  # warm-up
    g.clear_accum()
    t0 = time.perf_counter()
    g.render(spp, seed)
    wall = time.perf_counter() - t0
    st = g.stats()
    res = {"config": name, "width": w, "height": h, "spp": spp, "max_depth": depth, "seed": seed,
           "triangles": st["n_triangles"], "prims": st["n_prims"], "lights": st["n_lights"],
           "gpu": {"bvh_build_ms": build, "render_ms": st["ms_render"], "wall_s": wall,
                   "mrays_per_s": rays(st) / st["ms_render"] / 1e3, "mshadow_per_s": st["rays_shadow"] / st["ms_render"] / 1e3,
                   "spp_per_s": spp / (st["ms_render"] * 1e-3), "rays": rays(st), "shadow_rays": st["rays_shadow"],
                   "stage_ms": {k: st[k] for k in ("ms_generate", "ms_extend", "ms_shade", "ms_shadow", "ms_accumulate")}}}
    if save:
        host.write_image(save, host.accum_to_rgb8(g.read_accum(), spp))
    if orc_spp:
        o = oracle.context()
        sc.upload(ORC, o, w, h, depth)
        o.build_accel()
        t0 = time.perf_counter()
        o.render(orc_spp, seed)
        sec = time.perf_counter() - t0
        so = o.stats()
        # parity at the oracle's spp with equal seeds
        g.clear_accum()
        g.render(orc_spp, seed)
        a, b = g.read_accum() / orc_spp, o.read_accum() / orc_spp
        rmse = float(np.sqrt(np.mean((np.clip(a, 0, 1) - np.clip(b, 0, 1)) ** 2)))
        qa, qb = host.accum_to_rgb8(a, 1).astype(int), host.accum_to_rgb8(b, 1).astype(int)
        res["cpu_oracle"] = {"cores": CORES, "spp": orc_spp, "seconds": sec, "mrays_per_s": rays(so) / sec / 1e6,
                             "mshadow_per_s": so["rays_shadow"] / sec / 1e6}
        res["parity"] = {"spp": orc_spp, "rmse": rmse, "pixels_within_1_of_255": float(np.mean(np.all(np.abs(qa - qb) <= 1, axis=2))),
                         "ray_count_equal": rays(g.stats()) == rays(so), "shadow_count_equal": g.stats()["rays_shadow"] == so["rays_shadow"]}
    print(json.dumps(res), flush=True)
    return res


def soup_case(n_tris, out):
    import torch
    sc = host.Scene.builtin("soup", n_tris)
    g = mox.gpu().context(0)
    sc.upload(GPU, g, 64, 64, 5)
    builds = [g.build_accel() for _ in range(3)]
    glb = mox.gpu().context(0)
    sc.upload(GPU, glb, 64, 64, 5)
    lbvh = [glb.build_accel(mox.structs.ACCEL_LBVH) for _ in range(3)]
    res = {"config": f"soup {n_tris} triangles", "triangles": n_tris, "bvh_build_ms_ploc": min(builds), "bvh_build_ms_lbvh": min(lbvh),
           "build_bytes_per_tri": 450, "sweep": []}
    gen = torch.Generator(device="cuda").manual_seed(12345)
    for logn in range(20, 27):
        n = 1 << logn
        o = torch.rand((n, 3), generator=gen, device="cuda")
        d = torch.randn((n, 3), generator=gen, device="cuda")
        d = d / d.norm(dim=1, keepdim=True)
        r = torch.empty((n, 8), device="cuda")
        r[:, 0:3] = o; r[:, 3] = 1e-3; r[:, 4:7] = d; r[:, 7] = 1e27
        hits = torch.empty((n, 4), device="cuda")
        torch.cuda.synchronize()
        ms = min(g.trace_closest_device(r.data_ptr(), n, hits.data_ptr()) for _ in range(3))
        ms_l = min(glb.trace_closest_device(r.data_ptr(), n, hits.data_ptr()) for _ in range(3))
        hit_frac = float((hits[:, 1].view(torch.int32) >= 0).float().mean())
        res["sweep"].append({"rays": n, "ms_ploc": ms, "mrays_per_s_ploc": n / ms / 1e3, "ms_lbvh": ms_l,
                             "mrays_per_s_lbvh": n / ms_l / 1e3, "hit_fraction": hit_frac})
        del o, d, r, hits
    # roofline of the incoherent-ray batch: n-bar from a counting build of the same kernel on 2^24 rays
    gc = mox.gpu().context(0)
    sc.upload(GPU, gc, 64, 64, 5)
    gc.build_accel(mox.structs.ACCEL_COUNTERS)
    n = 1 << 24
    o = torch.rand((n, 3), generator=gen, device="cuda")
    d = torch.randn((n, 3), generator=gen, device="cuda")
    d = d / d.norm(dim=1, keepdim=True)
    r = torch.empty((n, 8), device="cuda")
    r[:, 0:3] = o; r[:, 3] = 1e-3; r[:, 4:7] = d; r[:, 7] = 1e27
    hits = torch.empty((n, 4), device="cuda")
    torch.cuda.synchronize()
    gc.trace_closest_device(r.data_ptr(), n, hits.data_ptr())
    cs = gc.stats()
    ms = min(g.trace_closest_device(r.data_ptr(), n, hits.data_ptr()) for _ in range(3))
    nn, npr = cs["node_visits"] / n, cs["prim_tests"] / n
    b_ray = 48 + nn * cs["node_bytes"] + npr * cs["prim_bytes"]
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    ach = n * b_ray / (ms * 1e-3) / 1e9
    res["roofline"] = {"bound": "hbm", "kernel": "k_traverse_wide<closest> (raw query, 2^24 incoherent rays)", "rays": n, "ms": ms,
                       "mrays_per_s": n / ms / 1e3, "nodes_per_ray": nn, "prims_per_ray": npr, "node_bytes": cs["node_bytes"],
                       "prim_bytes": cs["prim_bytes"], "bytes_per_ray": b_ray, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                       "scene_bytes": cs["n_nodes"] * cs["node_bytes"] + n_tris * cs["prim_bytes"]}
    res["roofline_build"] = {"bytes_per_triangle": 450, "ms": min(builds), "achieved": 450.0 * n_tris / (min(builds) * 1e-3) / 1e9,
                             "frac": 450.0 * n_tris / (min(builds) * 1e-3) / 1e9 / peak, "ms_lbvh": min(lbvh),
                             "frac_lbvh": 450.0 * n_tris / (min(lbvh) * 1e-3) / 1e9 / peak}
    print(json.dumps(res), flush=True)
    out.append(res)


def main():
    quick = "--quick" in sys.argv
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None   # e.g. --only C2
    if only == "C2":
        r = render_case("C2 random spheres 1920x1080 64spp thin lens", host.Scene.builtin("random_spheres"), 1920, 1080, 64, 5, 0x5EED, 0)
        print(json.dumps({"mrays_per_s": r["gpu"]["mrays_per_s"], "stage_ms": r["gpu"]["stage_ms"]}))
        return
    out = []
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    cornell = host.Scene.load(os.path.join(ROOT, "scenes", "cornell"), "cornell")
    out.append(render_case("C1 cornell 512x512 16spp", cornell, 512, 512, 16, 5, 0xC0FFEE, 16, save=os.path.join(ROOT, "gpurun_out", "c1_cornell.png")))
    out.append(render_case("C2 random spheres 1920x1080 64spp thin lens", host.Scene.builtin("random_spheres"), 1920, 1080, 64, 5, 0x5EED, 2,
                           save=os.path.join(ROOT, "gpurun_out", "c2_random_spheres.png")))
    if os.path.exists(os.path.join(ROOT, "scenes", "coffee", "coffee.scene")):
        coffee = host.Scene.load(os.path.join(ROOT, "scenes", "coffee"), "coffee")
        out.append(render_case("C3 coffee 1920x1080 256spp (Mesh010 missing upstream)", coffee, 1920, 1080, 64 if quick else 256, 5, 0xC0FFEE, 1,
                               save=os.path.join(ROOT, "gpurun_out", "c3_coffee.png")))
    interior = host.Scene.builtin("interior", 1000000)
    out.append(render_case("C4 interior ~1M tris 3840x2160", interior, 3840, 2160, 8 if quick else 64, 5, 0xD1A1A6, 0,
                           save=os.path.join(ROOT, "gpurun_out", "c4_interior.png")))
    soup_case(1000000 if quick else 10000000, out)
    json.dump({"host_cores": CORES, "results": out}, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
