import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import minimaloptix_b200 as mox, oracle
from minimaloptix_b200 import host
from conftest import random_rays
sc = host.Scene.builtin("interior", 30000)
o = oracle.context()
sc.upload(host.ApiTable(oracle.ORACLE_LIB, "orc_"), o, 64, 64, 5); o.build_accel()
rays = random_rays(100000, [0.05, 0.05, 0.05], [9.95, 3.95, 7.95], 8, tmax=3.0)
ao = o.trace_shadow(rays)
ob = oracle.context(brute_force=True)
sc.upload(host.ApiTable(oracle.ORACLE_LIB, "orc_"), ob, 64, 64, 5); ob.build_accel()
ab = ob.trace_shadow(rays[:20000])
print("oracle bvh vs brute mismatches (first 20k):", int(np.any(ab != ao[:20000], axis=1).sum()))
api = host.ApiTable(mox.GPU_LIB, "mox_")
for it in range(4):
    g = mox.gpu().context(0)
    sc.upload(api, g, 64, 64, 5); g.build_accel()
    ag = g.trace_shadow(rays)
    bad = np.any(ao != ag, axis=1)
    zero_o, zero_g = (ao == 0).all(axis=1), (ag == 0).all(axis=1)
    print(it, "mismatch", int(bad.sum()), "oracle blocked/gpu not", int((zero_o & ~zero_g).sum()), "gpu blocked/oracle not", int((~zero_o & zero_g).sum()),
          "max abs diff among both-unblocked", float(np.abs(ao - ag)[bad & ~zero_o & ~zero_g].max()) if (bad & ~zero_o & ~zero_g).any() else 0.0)
    k = np.nonzero(bad)[0][:3]
    for i in k: print("   ray", i, ao[i], ag[i])
