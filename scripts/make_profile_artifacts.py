#!/usr/bin/env python3
"""Turn the captures of scripts/gpu_profile_r2.sh (gpurun_out/launches_SRC.csv, gpurun_out/prof_SRC.ncu-rep) into
the committed evidence under profiles/: launch list, per-kernel share of the step, counter summary of the hot
kernels, and the DRAM + L2 traffic per launch that bench.py reports as roofline.traffic / l2_bytes_per_launch.
usage: make_profile_artifacts.py TAG [SRC] [SPP]   (files are written as profiles/TAG_*; SRC defaults to TAG,
SPP = samples per wavefront of the captured launches, default 4)"""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
tag = sys.argv[1]
src_tag = sys.argv[2] if len(sys.argv) > 2 else tag
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 4
out = lambda name: os.path.join(ROOT, "profiles", f"{tag}_{name}")

# ---- launch list -> shares
src = os.path.join(ROOT, "gpurun_out", f"launches_{src_tag}.csv")
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

# ---- full captures (raw-page CSVs of the traversal and of the shading kernels) -> summary + traffic
summary = ""
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tscale = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}
key_of = lambda n: ("k_traverse_closest" if "k_traverse_wide<0" in n else "k_traverse_shadow" if "k_traverse_wide<1" in n
                    else "k_shade_disney" if "k_shade_disney" in n else "k_classify" if "k_classify" in n else "k_logic" if "k_logic" in n
                    else "k_apply" if "k_apply" in n else "k_accumulate" if "k_accumulate" in n else "k_generate" if "k_generate" in n else None)
acc = {}
for part in ("trav", "shade"):
    f = os.path.join(ROOT, "gpurun_out", f"prof_{src_tag}_{part}_raw.csv")
    if not os.path.exists(f):
        continue
    summary += f"== {part}: ncu --set full --metrics lts__t_bytes.sum --clock-control none, launches of one step, {spp} spp per wavefront ==\n"
    summary += subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), f], capture_output=True, text=True).stdout
    rows_ = [r for r in csv.reader(open(f)) if r]
    st = [i for i, r in enumerate(rows_) if r[0] == "ID"][0]
    h, units, data = rows_[st], rows_[st + 1], rows_[st + 2:]   # units are per column AND per file
    col = {n: h.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "lts__t_bytes.sum") if n in h}
    val = lambda r, c: float(r[col[c]]) * scale[units[col[c]]]
    for r in data:
        k = key_of(r[col["Kernel Name"]])
        if not k:
            continue
        a = acc.setdefault(k, {"kernel": r[col["Kernel Name"]][:64], "launches_profiled": 0, "dram_bytes_per_launch": 0.0,
                               "lts_bytes_per_launch": 0.0, "ms_per_launch_under_ncu": 0.0})
        a["launches_profiled"] += 1
        a["dram_bytes_per_launch"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        if "lts__t_bytes.sum" in col:
            a["lts_bytes_per_launch"] += val(r, "lts__t_bytes.sum")
        a["ms_per_launch_under_ncu"] += float(r[col["gpu__time_duration.sum"]]) * tscale.get(units[col["gpu__time_duration.sum"]], 1.0)
open(out("kernels_summary.txt"), "w").write(summary)
for a in acc.values():
    for k in ("dram_bytes_per_launch", "lts_bytes_per_launch", "ms_per_launch_under_ncu"):
        a[k] /= a["launches_profiled"]
    a["dram_gbs_under_ncu"] = a["dram_bytes_per_launch"] / (a["ms_per_launch_under_ncu"] * 1e-3) / 1e9
    a["lts_gbs_under_ncu"] = a["lts_bytes_per_launch"] / (a["ms_per_launch_under_ncu"] * 1e-3) / 1e9
acc["_source"] = (f"ncu --set full capture (summary: profiles/{tag}_kernels_summary.txt): launches of one 4K step of the "
                  f"bench scene, {spp} spp per wavefront, 8-wide BVH; dram = dram__bytes_read.sum + dram__bytes_write.sum, "
                  "lts = lts__t_bytes.sum, averaged over the profiled launches of each kernel")
json.dump(acc, open(os.path.join(ROOT, "profiles", f"{tag.split('_')[0]}_traffic.json"), "w"), indent=1)
print(json.dumps(acc, indent=1))
