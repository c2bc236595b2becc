#!/usr/bin/env python3
"""Turn the captures of scripts/gpu_profile_final.sh (gpurun_out/launches_final.csv, gpurun_out/prof_final.ncu-rep)
into the committed evidence under profiles/: launch list, per-kernel share of the step, counter summary of
the hot kernels and the DRAM traffic per launch that bench.py reports as roofline.traffic.
usage: make_profile_artifacts.py TAG   (files are written as profiles/TAG_*)"""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
tag = sys.argv[1]
out = lambda name: os.path.join(ROOT, "profiles", f"{tag}_{name}")

# ---- launch list -> shares
src = os.path.join(ROOT, "gpurun_out", "launches_final.csv")
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += float(r[vi].replace(",", "")) / 1e3
shutil.copy(src, out("launches.csv"))
total = sum(v[1] for v in tot.values())
with open(out("launch_shares.csv"), "w") as f:
    f.write("kernel,launches,total_us,share_of_listed_gpu_time\n")
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n},{us:.1f},{us / total:.4f}\n")
print(open(out("launch_shares.csv")).read())

# ---- full capture -> summary + traffic
rep = os.path.join(ROOT, "gpurun_out", "prof_final.ncu-rep")
summary = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
open(out("kernels_summary.txt"), "w").write(summary)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units, data = rr[0], rr[1], rr[2:]
col = {n: h.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
key_of = lambda n: ("k_traverse_closest" if "k_traverse_wide<0" in n else "k_traverse_shadow" if "k_traverse_wide<1" in n
                    else "k_shade_disney" if "k_shade_disney" in n else "k_logic" if "k_logic" in n else "k_apply" if "k_apply" in n else None)
acc = {}
for r in data:
    k = key_of(r[col["Kernel Name"]])
    if not k:
        continue
    b = sum(float(r[col[c]]) * scale[units[col[c]]] for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    a = acc.setdefault(k, {"kernel": r[col["Kernel Name"]][:64], "launches_profiled": 0, "dram_bytes_per_launch": 0.0, "ms_per_launch_under_ncu": 0.0})
    a["launches_profiled"] += 1
    a["dram_bytes_per_launch"] += b
    a["ms_per_launch_under_ncu"] += float(r[col["gpu__time_duration.sum"]])
for a in acc.values():
    a["dram_bytes_per_launch"] /= a["launches_profiled"]
    a["ms_per_launch_under_ncu"] /= a["launches_profiled"]
acc["_source"] = (f"ncu --set full capture (summary: profiles/{tag}_kernels_summary.txt): bounce launches of one 4K step of the "
                  "bench scene, 1 spp per wavefront, 8-wide BVH")
json.dump(acc, open(os.path.join(ROOT, "profiles", "r1_final_traffic.json"), "w"), indent=1)
print(json.dumps(acc, indent=1))
