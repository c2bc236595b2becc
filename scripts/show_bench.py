#!/usr/bin/env python3
import json, sys
d = json.load(open(sys.argv[1]))
for k in ("roofline_closest", "roofline_shadow"):
    r = d.get(k)
    if r:
        print(k, {x: (round(r[x], 3) if isinstance(r[x], float) else r[x]) for x in ("achieved", "frac", "traffic", "bytes_per_ray", "nodes_per_ray", "prims_per_ray", "rays_per_s", "share_of_step")})
print("value", d["value"], "e2e", d["e2e"], "cpu", d.get("cpu_baseline"), "stage", d["stage_ms"])
