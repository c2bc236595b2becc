#!/usr/bin/env python3
"""the animated random-spheres sequence as a rebuild benchmark — per frame:
animate(0.002) -> update_sphere x259 -> mox_build_accel -> 4 spp at 1920x1080."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import minimaloptix_b200 as mox
from minimaloptix_b200 import host

def main():
    frames, spp = 60, 4
    api = host.ApiTable(mox.GPU_LIB, "mox_")
    sc = host.Scene.builtin("random_spheres")
    g = mox.gpu().context(0)
    sc.upload(api, g, 1920, 1080, 5)
    g.build_accel(); g.render(1, 1); g.clear_accum()
    builds, renders = [], []
    t0 = time.perf_counter()
    for f in range(frames):
        sc.animate(0.002)
        sc.apply_spheres(api, g)
        builds.append(g.build_accel())
        g.clear_accum()
        g.render(spp, 1000 + f)
        renders.append(g.stats()["ms_render"])
        g.map_accum()
    wall = time.perf_counter() - t0
    print(json.dumps({"frames": frames, "spp": spp, "bvh_rebuild_ms_avg": sum(builds) / frames, "bvh_rebuild_ms_max": max(builds),
                      "render_ms_avg": sum(renders) / frames, "fps_wall": frames / wall}))
if __name__ == "__main__":
    main()
