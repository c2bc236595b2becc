#!/usr/bin/env python3
"""ncu target: build the acceleration structure of a built-in scene twice (the second build reuses the arena).
usage: build_only.py [interior|soup] [triangles]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import minimaloptix_b200 as mox
from minimaloptix_b200 import host
name = sys.argv[1] if len(sys.argv) > 1 else "interior"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
sc = host.Scene.builtin(name, n)
g = mox.gpu().context(0)
sc.upload(host.ApiTable(mox.GPU_LIB, "mox_"), g, 64, 64, 5)
print("build ms", [g.build_accel() for _ in range(2)])
