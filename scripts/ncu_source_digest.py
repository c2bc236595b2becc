#!/usr/bin/env python3
"""Digest of an ncu source-page CSV (ncu -i REP --page source --csv): per kernel launch the stall mix and the SASS
instructions that collect the most warp-stall samples.  usage: ncu_source_digest.py CSV > profiles/TAG_trav_source_top.txt"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
print("ncu --set full --import-source on, source page (SASS) of the two traversal kernels, bench scene, 4 spp per wavefront")
print("(first launch of each kernel in the capture window: camera rays, and the shadow rays of depth 1).  Per launch: the")
print("stall mix, and the instructions with the most warp-stall samples (share of all samples of the launch), their threads")
print("per executed instruction and dominant stall reason.  Made by scripts/gpu_profile_r2.sh + scripts/ncu_source_digest.py.\n")
seen = set()
for b in range(len(starts) - 1):
    blk = rows[starts[b]:starts[b + 1]]
    name, hdr = blk[0][1], blk[1]
    data = [r for r in blk[2:] if len(r) >= len(hdr) - 2]
    ci = {n: i for i, n in enumerate(hdr)}
    num = lambda r, c: int(r[ci[c]] or 0)
    tot = sum(num(r, "# Samples") for r in data)
    inst = sum(num(r, "Instructions Executed") for r in data)
    thr = sum(num(r, "Thread Instructions Executed") for r in data)
    key = (name, tot, inst)
    if key in seen:
        continue
    seen.add(key)
    stall_cols = [n for n in hdr if n.startswith("stall_") and "(Not Issued)" not in n]
    agg = {n: sum(num(r, n) for r in data) for n in stall_cols}
    print(f"=== {name}")
    print(f"    {len(data)} SASS instructions, {inst} warp instructions executed, {thr / max(inst, 1):.2f} threads per instruction, {tot} stall samples")
    print("    stall mix: " + ", ".join(f"{k[6:]} {v / tot:.1%}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:14]:
        main = max(stall_cols, key=lambda n: num(r, n))
        ie, te = num(r, "Instructions Executed"), num(r, "Thread Instructions Executed")
        print(f"    {num(r, '# Samples') / tot:6.2%}  {r[ci['Source']].strip()[:70]:70s} thr/inst {te / max(ie, 1):5.1f}  mostly {main[6:]}")
    print()
