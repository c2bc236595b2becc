#!/usr/bin/env python
"""Generate tests/golden/render_ref.npz (+ .json) from the REFERENCE'S OWN device code.

Runs only where /root/reference exists: `make ref` compiles Camera.cu, Geometry.cu, Material.cu,
miss.cu, Exception.cu, disney.h and utils_device.h unchanged into oracle/_ref/libref_render.so
(shim: oracle/ref_shim/render/optix_world.h, harness: oracle/ref_shim/ref_render.cpp); this script
evaluates tests/refcases.py on it and stores every result.  tests/test_ref_render.py then holds
oracle/ to these vectors bit for bit on any machine.

It also runs BASELINE config 1 at full size (Cornell 512x512, 16 spp, depth 5) on both the
compiled reference and the oracle and records the comparison in the .json.
"""
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refcases as R  # noqa: E402
from minimaloptix_b200 import host  # noqa: E402

REF_SRC = "/root/reference/MinimalOptiX"
SOURCES = ["Camera.cu", "Geometry.cu", "Material.cu", "miss.cu", "Exception.cu", "disney.h", "utils_device.h", "Structures.h"]


def main():
    subprocess.check_call(["make", "-C", ROOT, "ref", "oracle", "host"])
    ref = R.load("ref")
    t0 = time.time()
    res = R.all_cases(ref, host)
    np.savez_compressed(R.GOLDEN, **res)
    meta = {
        "generator": "scripts/make_render_golden.py",
        "library": "oracle/_ref/libref_render.so (reference device sources compiled unchanged with g++)",
        "compiler": subprocess.check_output(["g++", "--version"], text=True).splitlines()[0],
        "flags": "-O2 -std=c++14 -ffp-contract=off -fno-fast-math",
        "sources_sha256": {s: hashlib.sha256(open(os.path.join(REF_SRC, s), "rb").read()).hexdigest() for s in SOURCES},
        "draw_order_note": "g++ gives the first draw of Material.cu:180 to light.v (orc_set_quad_light_draw_order(1)); "
                           "all other in-expression draws are pinned left to right by the shim",
        "cases": {k: list(v.shape) for k, v in res.items()},
        "seconds": round(time.time() - t0, 2),
    }
    # BASELINE config 1 at full size, reference vs oracle, bit for bit
    orc = R.load("orc")
    t0 = time.time()
    a = R.scene_case(ref, host, "cornell", size=(512, 512), spp=16)
    t1 = time.time()
    b = R.scene_case(orc, host, "cornell", size=(512, 512), spp=16)
    t2 = time.time()
    meta["config1_cornell_512x512_16spp_depth5"] = {
        "image_bits_equal": bool(R.bits_equal(a["image"], b["image"])),
        "floats_differing": int((a["image"].view(np.uint32) != b["image"].view(np.uint32)).sum()),
        "ray_counts_reference": a["ray_counts"].tolist(), "ray_counts_oracle": b["ray_counts"].tolist(),
        "image_sha256": hashlib.sha256(a["image"].tobytes()).hexdigest(),
        "seconds_reference": round(t1 - t0, 1), "seconds_oracle": round(t2 - t1, 1),
    }
    with open(R.GOLDEN.replace(".npz", ".json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta["config1_cornell_512x512_16spp_depth5"], indent=1))
    print("wrote", R.GOLDEN, os.path.getsize(R.GOLDEN), "bytes")


if __name__ == "__main__":
    main()
