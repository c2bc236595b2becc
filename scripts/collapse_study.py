#!/usr/bin/env python3
"""CPU model of the wide collapse: how much SAH cost does the optimal collapse of the binary PLOC tree into the
8-wide BVH (bvh_wide.cu::k_collapse_dp) save over the greedy largest-area-first collapse it replaced, and what
would other leaf limits give?  (It predicted 4.5-5.4 % before the GPU version was written; the bench gained 4.2 %.)

Emulates the GPU builder in numpy on the bench scene at a reduced triangle count (Morton order -> PLOC with the
same radius -> binary tree), then collapses it twice:
  greedy   open the child with the largest surface area until there are 8 (subtrees of <= 2 primitives stay closed
           as leaf children; free slots are then filled by opening those too) — MOX_WIDE_GREEDY=1, > 4 M primitives;
  optimal  the dynamic programme of Ylitie, Karras & Laine 2017 (C(n, i), i = 1..7, C_distribute(n, 8)).
Cost model: c_node per visited wide node, c_prim per tested primitive, weighted by surface area (c_prim / c_node
= 0.43: the measured instruction counts of a primitive step and a node step).  No GPU needed.
usage: collapse_study.py [triangles=60000] [radius=32] [--topdown]"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from minimaloptix_b200 import host  # noqa: E402

C_NODE, C_PRIM, LEAF_MAX, WIDTH = 1.0, 0.43, 2, 8


def area(lo, hi):
    d = np.maximum(hi - lo, 0)
    return d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0]


def morton(c):
    q = np.clip((c * 1024).astype(np.int64), 0, 1023)

    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    return (spread(q[:, 0]) << 2) | (spread(q[:, 1]) << 1) | spread(q[:, 2])


def ploc(lo, hi, radius):
    """Returns (left, right, lo, hi, size) of a binary tree; nodes 0..n-1 are the primitives in Morton order."""
    n = len(lo)
    L, R = [-1] * n, [-1] * n
    nlo, nhi, size = [*lo], [*hi], [1] * n
    cid = np.arange(n)
    clo, chi = lo.copy(), hi.copy()
    while len(cid) > 1:
        m = len(cid)
        best = np.full(m, np.inf)
        nn = np.full(m, -1)
        for o in list(range(-radius, 0)) + list(range(1, radius + 1)):   # ascending j: first minimum wins
            j = np.arange(m) + o
            ok = (j >= 0) & (j < m)
            jj = np.clip(j, 0, m - 1)
            a = area(np.minimum(clo, clo[jj]), np.maximum(chi, chi[jj]))
            a[~ok] = np.inf
            better = a < best
            best[better] = a[better]
            nn[better] = jj[better]
        mutual = nn[nn] == np.arange(m)
        keep = np.ones(m, bool)
        out_cid, out_lo, out_hi = cid.copy(), clo.copy(), chi.copy()
        for i in np.nonzero(mutual & (np.arange(m) < nn))[0]:
            j = nn[i]
            node = len(L)
            L.append(int(cid[i])); R.append(int(cid[j]))
            b_lo, b_hi = np.minimum(clo[i], clo[j]), np.maximum(chi[i], chi[j])
            nlo.append(b_lo); nhi.append(b_hi); size.append(size[cid[i]] + size[cid[j]])
            out_cid[i], out_lo[i], out_hi[i] = node, b_lo, b_hi
            keep[j] = False
        cid, clo, chi = out_cid[keep], out_lo[keep], out_hi[keep]
    return np.array(L), np.array(R), np.array(nlo), np.array(nhi), np.array(size)


def topdown_sah(lo, hi, bins=16):
    """Binned-SAH top-down binary tree (the classic CPU quality builder) over the same primitives: an estimate of
    how much tree quality is left beyond PLOC.  Nodes are appended in post-order (children before parents)."""
    sys.setrecursionlimit(100000)
    n = len(lo)
    L, R = [-1] * n, [-1] * n
    nlo, nhi, size = [*lo], [*hi], [1] * n
    cen = 0.5 * (lo + hi)

    def build(ids):
        if len(ids) == 1:
            return int(ids[0])
        blo, bhi = lo[ids].min(axis=0), hi[ids].max(axis=0)
        best = (np.inf, None)
        c = cen[ids]
        cmin, cmax = c.min(axis=0), c.max(axis=0)
        for ax in range(3):
            if cmax[ax] <= cmin[ax]:
                continue
            b = np.minimum(((c[:, ax] - cmin[ax]) / (cmax[ax] - cmin[ax]) * bins).astype(int), bins - 1)
            for split in range(1, bins):
                left = b < split
                nl = int(left.sum())
                if nl == 0 or nl == len(ids):
                    continue
                il, ir = ids[left], ids[~left]
                cost = area(lo[il].min(axis=0), hi[il].max(axis=0)) * nl + area(lo[ir].min(axis=0), hi[ir].max(axis=0)) * (len(ids) - nl)
                if cost < best[0]:
                    best = (cost, left)
        if best[1] is None:
            half = len(ids) // 2
            il, ir = ids[:half], ids[half:]
        else:
            il, ir = ids[best[1]], ids[~best[1]]
        a, b2 = build(il), build(ir)
        L.append(a); R.append(b2)
        nlo.append(blo); nhi.append(bhi); size.append(len(ids))
        return len(L) - 1

    build(np.arange(n))
    return np.array(L), np.array(R), np.array(nlo), np.array(nhi), np.array(size)


def greedy_cost(L, R, A, size, root):
    total_nodes = total_leaf = 0.0
    count = 0
    stack = [root]
    fill = []
    while stack:
        n = stack.pop()
        count += 1
        total_nodes += A[n]
        ch = [L[n], R[n]]
        while len(ch) < WIDTH:
            openable = [c for c in ch if size[c] > LEAF_MAX]
            if not openable:
                break
            c = max(openable, key=lambda x: A[x])
            ch.remove(c); ch += [L[c], R[c]]
        while len(ch) < WIDTH:   # fill free slots with single primitives
            two = [c for c in ch if size[c] == 2]
            if not two:
                break
            c = max(two, key=lambda x: A[x])
            ch.remove(c); ch += [L[c], R[c]]
        fill.append(len(ch))
        for c in ch:
            if size[c] > LEAF_MAX:
                stack.append(c)
            else:
                total_leaf += A[c] * size[c]
    return C_NODE * total_nodes + C_PRIM * total_leaf, count, float(np.mean(fill))


def optimal_cost(L, R, A, size, root, n_prims):
    N = len(L)
    C = np.full((N, WIDTH), np.inf)          # C[n, i] = cost of node n as a forest of at most i roots (i = 1..7 used)
    internal = np.zeros(N, bool)
    for n in range(N):                        # children always have smaller ids than their parent
        if n < n_prims:
            C[n, 1:] = A[n] * C_PRIM
            continue
        l, r = L[n], R[n]
        dist = np.full(WIDTH + 1, np.inf)
        for j in range(2, WIDTH + 1):
            for k in range(1, j):
                if k < WIDTH and j - k < WIDTH:
                    dist[j] = min(dist[j], C[l, k] + C[r, j - k])
        c_leaf = A[n] * size[n] * C_PRIM if size[n] <= LEAF_MAX else np.inf
        c_int = dist[WIDTH] + A[n] * C_NODE
        C[n, 1] = min(c_leaf, c_int)
        internal[n] = c_int < c_leaf
        for i in range(2, WIDTH):
            C[n, i] = min(dist[i], C[n, i - 1])
    return C[root, 1]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_tris = int(args[0]) if len(args) > 0 else 60000
    radius = int(args[1]) if len(args) > 1 else 32
    sc = host.Scene.builtin("interior", n_tris)
    tris = []
    for m in range(sc.info().n_meshes):
        v, idx = sc.mesh_arrays(m)
        tris.append(v[idx])
    t = np.concatenate(tris).astype(np.float64)
    lo, hi = t.min(axis=1), t.max(axis=1)
    cen = 0.5 * (lo + hi)
    smin, smax = cen.min(axis=0), cen.max(axis=0)
    order = np.argsort(morton((cen - smin) / np.maximum(smax - smin, 1e-30)), kind="stable")
    lo, hi = lo[order], hi[order]
    L, R, nlo, nhi, size = ploc(lo, hi, radius)
    root = len(L) - 1
    A = area(nlo, nhi) / area(nlo[root], nhi[root])
    g, n_wide, fill = greedy_cost(L, R, A, size, root)
    o = optimal_cost(L, R, A, size, root, len(lo))
    print(f"{len(lo)} triangles, PLOC radius {radius}: binary nodes {len(L) - len(lo)}")
    print(f"greedy collapse : SAH cost {g:.3f}  ({n_wide} wide nodes, {fill:.2f} children per node)")
    print(f"optimal collapse: SAH cost {o:.3f}  ({100 * (1 - o / g):.1f} % lower)")
    if "--topdown" in sys.argv:
        L2, R2, nlo2, nhi2, size2 = topdown_sah(lo, hi)
        root2 = len(L2) - 1
        A2 = area(nlo2, nhi2) / area(nlo2[root2], nhi2[root2])
        o2 = optimal_cost(L2, R2, A2, size2, root2, len(lo))
        print(f"binned-SAH top-down tree, optimal collapse: SAH cost {o2:.3f}  ({100 * (1 - o2 / o):.1f} % below PLOC + optimal collapse)")
    global LEAF_MAX
    for lm in (1, 3, 4):   # what the programme would make of other leaf limits (the shipped limit is 2)
        LEAF_MAX = lm
        print(f"optimal collapse with leaf children of <= {lm} primitives: SAH cost {optimal_cost(L, R, A, size, root, len(lo)):.3f}")


if __name__ == "__main__":
    main()
