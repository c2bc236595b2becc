#!/usr/bin/env python3
"""Instruction histograms of the hot kernels from the built library (cuobjdump -sass; no GPU needed): whole kernel and,
for the wide traversal kernels, the node step alone (from the pop of the front-most child to the branch that ends the
step).  usage: sass_histogram.py [lib] > profiles/TAG_sass_histogram.txt"""
import collections, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else "minimaloptix_b200/libmox.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), [])
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*;", line)
    if m and cur is not None:
        cur.append(m.group(1))


def opcode(ins):
    ins = re.sub(r"^@!?U?P\w+\s+", "", ins)
    return ins.split()[0].split(".")[0]


def hist(instrs):
    c = collections.Counter(opcode(i) for i in instrs)
    return ", ".join(f"{k} {v}" for k, v in c.most_common())


want = [("k_traverse_wide<closest, render path>", "15k_traverse_wideILb0ELb0ELb1E"),
        ("k_traverse_wide<anyhit> (shadow rays)", "15k_traverse_wideILb1ELb0ELb0E"),
        ("k_shade_disney<ref rng, fast BRDF values>", "14k_shade_disneyILi0ELb1E"),
        ("k_classify", "10k_classify"), ("k_apply", "7k_apply")]
print(f"cuobjdump -sass {lib} (sm_100a), made by scripts/sass_histogram.py\n")
for title, pat in want:
    for name, ins in kernels.items():
        if pat in name:
            print(f"== {title}: {len(ins)} instructions\n   {hist(ins)}")
            if "traverse_wide" in pat:
                # node step: the four-or-five 128-bit loads of the 80-byte node mark it; it starts at the FLO before them
                loads = [i for i, x in enumerate(ins) if x.startswith("LDG.E.128.CONSTANT") and "+0x40]" in x]
                if loads:
                    a = max(i for i in range(loads[0]) if ins[i].startswith("FLO"))
                    b = next(i for i in range(loads[0] + 100, len(ins)) if ins[i].startswith("BRA"))
                    step = ins[a:b]
                    print(f"   node step: {len(step)} instructions\n   {hist(step)}")
            print()
            break
