#!/usr/bin/env python3
"""CPU model for the next round: does the order in which a shadow ray visits the children of a wide node matter?

Shadow rays are 40 % of a bench step.  They stop at the first blocker, so for occluded rays the visiting order decides
how many nodes are fetched; the kernel uses the closest-hit order (front to back by octant).  This script builds the
8-wide BVH as the GPU does (numpy PLOC + the optimal collapse of collapse_study.py, slots by centroid octant), shoots
shadow rays from random surface points to random points on the scene's quad lights and counts node visits and
primitive tests per ray for several orderings of the hit inner children:
  octant    slot ^ ray octant, highest first (what traverse_wide.cuh does)
  near      entry distance, nearest first
  area      surface area, largest first
  overlap   length of the ray segment inside the box, longest first
  size      primitives below, most first
Result (15 k triangles, 1 500 rays, 33 % occluded): octant 4.64 node visits per ray, overlap 4.66, area 4.81, near 4.86,
size 5.03 — the order the kernel uses is already the best of these.  With --closest the same model follows bounce-like
rays to their closest hit: octant order costs 7.18 node visits per ray against 7.08 for exact nearest-first, so the order
is fine; what the kernel's stack lacks is distances — dropping a postponed group whose nearest member lies beyond the
current best hit would save 7 % of the node visits (6.66), per-member distances 8.5 % (6.57).
No GPU needed.  usage: shadow_order_study.py [triangles=40000] [rays=4000] [--closest]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import collapse_study as cs  # noqa: E402
from minimaloptix_b200 import host  # noqa: E402

GLASS = 1


def decisions(L, R, A, size, n_prims):
    """The collapse programme with its argmins kept: per node C[1..7], best k per j = 2..8, inherit flags."""
    N, W = len(L), cs.WIDTH
    C = np.full((N, W), np.inf)
    kbest = np.zeros((N, W + 1), int)
    inherit = np.zeros((N, W), bool)
    for n in range(N):
        if n < n_prims:
            C[n, 1:] = A[n] * cs.C_PRIM
            continue
        l, r = L[n], R[n]
        dist = np.full(W + 1, np.inf)
        for j in range(2, W + 1):
            for k in range(1, j):
                if k < W and j - k < W and C[l, k] + C[r, j - k] < dist[j]:
                    dist[j] = C[l, k] + C[r, j - k]
                    kbest[n, j] = k
        c_leaf = A[n] * size[n] * cs.C_PRIM if size[n] <= cs.LEAF_MAX else np.inf
        C[n, 1] = min(c_leaf, dist[W] + A[n] * cs.C_NODE)
        for i in range(2, W):
            if dist[i] < C[n, i - 1]:
                C[n, i] = dist[i]
            else:
                C[n, i] = C[n, i - 1]
                inherit[n, i] = True
    return kbest, inherit


def collapse(L, R, size, n_prims, root, kbest, inherit):
    """Wide nodes: list of children (binary node ids) per wide node; wide index of every inner child."""
    wide, index_of, queue = [], {root: 0}, [root]
    while queue:
        n = queue.pop(0)
        ch, st = [], [(R[n], cs.WIDTH - kbest[n, cs.WIDTH]), (L[n], kbest[n, cs.WIDTH])]
        while st:
            m, i = st.pop()
            if m < n_prims:
                ch.append(m)
                continue
            while i > 1 and inherit[m, i]:
                i -= 1
            if i <= 1:
                ch.append(m)
                continue
            st += [(R[m], i - kbest[m, i]), (L[m], kbest[m, i])]
        wide.append(ch)
        for c in ch:
            if size[c] > cs.LEAF_MAX:
                index_of[c] = len(index_of)
                queue.append(c)
    return wide, index_of


def leaves_of(L, R, n_prims, node):
    out, st = [], [node]
    while st:
        m = st.pop()
        if m < n_prims:
            out.append(m)
        else:
            st += [L[m], R[m]]
    return out


def tri_hit(o, d, tmin, tmax, p0, p1, p2):
    e0, e1 = p1 - p0, p0 - p2
    n = np.cross(e1, e0)
    den = np.dot(n, d)
    if den == 0:
        return False
    e2 = (p0 - o) / den
    i = np.cross(d, e2)
    beta, gamma, t = np.dot(i, e1), np.dot(i, e0), np.dot(n, e2)
    return tmin < t < tmax and beta >= 0 and gamma >= 0 and beta + gamma <= 1


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_tris = int(args[0]) if args else 40000
    n_rays = int(args[1]) if len(args) > 1 else 4000
    rng = np.random.default_rng(7)
    sc = host.Scene.builtin("interior", n_tris)
    tris, blocks = [], []
    for m in range(sc.info().n_meshes):
        v, idx = sc.mesh_arrays(m)
        tris.append(v[idx])
        blocks.append(np.full(len(idx), sc.mesh_info(m)["disney"].brdfType != GLASS))
    t = np.concatenate(tris).astype(np.float64)
    blocks = np.concatenate(blocks)
    lo, hi = t.min(axis=1), t.max(axis=1)
    cen = 0.5 * (lo + hi)
    smin, smax = cen.min(axis=0), cen.max(axis=0)
    order = np.argsort(cs.morton((cen - smin) / np.maximum(smax - smin, 1e-30)), kind="stable")
    t, blocks, lo, hi = t[order], blocks[order], lo[order], hi[order]
    n = len(t)
    L, R, nlo, nhi, size = cs.ploc(lo, hi, 32)
    root = len(L) - 1
    A = cs.area(nlo, nhi) / cs.area(nlo[root], nhi[root])
    kbest, inherit = decisions(L, R, A, size, n)
    wide, index_of = collapse(L, R, size, n, root, kbest, inherit)
    # per wide node: child boxes, kind, payload, slot (centroid octant, nearest free slot by Hamming distance)
    nodes = []
    for wi, ch in enumerate(wide):
        bn = [k for k, v in index_of.items() if v == wi][0] if wi else root
        c0 = 0.5 * (nlo[bn] + nhi[bn])
        used, entries = set(), []
        for c in ch:
            cc = 0.5 * (nlo[c] + nhi[c])
            want = (4 if cc[0] > c0[0] else 0) | (2 if cc[1] > c0[1] else 0) | (1 if cc[2] > c0[2] else 0)
            slot = min((s for s in range(8) if s not in used), key=lambda s: bin(s ^ want).count("1"))
            used.add(slot)
            inner = size[c] > cs.LEAF_MAX
            entries.append((nlo[c], nhi[c], inner, index_of[c] if inner else leaves_of(L, R, n, c), slot, A[c], size[c]))
        nodes.append(entries)
    # shadow rays: surface point (area-weighted) -> point on a quad light
    lights = [sc.light(i) for i in range(sc.info().n_lights)]
    tri_area = 0.5 * np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1)
    pick = rng.choice(n, size=4 * n_rays, p=tri_area / tri_area.sum())
    rays = []
    for ti in pick:
        u, v = rng.random(2)
        if u + v > 1:
            u, v = 1 - u, 1 - v
        p = t[ti, 0] + u * (t[ti, 1] - t[ti, 0]) + v * (t[ti, 2] - t[ti, 0])
        ng = np.cross(t[ti, 1] - t[ti, 0], t[ti, 2] - t[ti, 0])
        ng /= np.linalg.norm(ng) + 1e-30
        lp = lights[rng.integers(len(lights))]
        q = np.array(lp.position.tuple()) + rng.random() * np.array(lp.u.tuple()) + rng.random() * np.array(lp.v.tuple())
        d = q - p
        dist = np.linalg.norm(d)
        d /= dist
        nl = np.array(lp.normal.tuple())
        if np.dot(d, nl) >= 0:
            continue
        side = 1.0 if np.dot(d, ng) > 0 else -1.0
        rays.append((p + side * 1e-4 * ng, d, 1e-4, dist - 1e-4))
        if len(rays) == n_rays:
            break

    def trace(o, d, tmin, tmax, policy):
        inv = 1.0 / np.where(np.abs(d) > 1e-30, d, 1e-30)
        octinv = 7 ^ ((4 if d[0] < 0 else 0) | (2 if d[1] < 0 else 0) | (1 if d[2] < 0 else 0))
        visits = tests = 0
        stack = [0]
        while stack:
            wi = stack.pop()
            visits += 1
            inner = []
            for (blo, bhi, is_inner, payload, slot, a, sz) in nodes[wi]:
                t0, t1 = (blo - o) * inv, (bhi - o) * inv
                tn, tf = max(np.minimum(t0, t1).max(), tmin), min(np.maximum(t0, t1).min(), tmax)
                if tn > tf:
                    continue
                if is_inner:
                    key = {"octant": slot ^ octinv, "near": -tn, "area": a, "overlap": tf - tn, "size": sz}[policy]
                    inner.append((key, payload))
                else:
                    for prim in payload:
                        tests += 1
                        if blocks[prim] and tri_hit(o, d, tmin, tmax, t[prim, 0], t[prim, 1], t[prim, 2]):
                            return visits, tests, True
            for key, payload in sorted(inner, key=lambda e: e[0]):   # last pushed = highest key = visited first
                stack.append(payload)
        return visits, tests, False

    def tri_t(o, d, tmin, tmax, p0, p1, p2):
        e0, e1 = p1 - p0, p0 - p2
        nn = np.cross(e1, e0)
        den = np.dot(nn, d)
        if den == 0:
            return None
        e2 = (p0 - o) / den
        i = np.cross(d, e2)
        beta, gamma, tt = np.dot(i, e1), np.dot(i, e0), np.dot(nn, e2)
        return tt if (tmin < tt < tmax and beta >= 0 and gamma >= 0 and beta + gamma <= 1) else None

    def trace_closest(o, d, tmin, policy):
        """Closest hit with the kernel's schedule: the primitives of a node's hit leaf children first, then its hit
        inner children in `policy` order, depth first; a postponed child is culled against the current best when popped
        only if `policy` keeps entry distances (the kernel's stack holds none: octant order never re-tests)."""
        inv = 1.0 / np.where(np.abs(d) > 1e-30, d, 1e-30)
        octinv = 7 ^ ((4 if d[0] < 0 else 0) | (2 if d[1] < 0 else 0) | (1 if d[2] < 0 else 0))
        visits = tests = 0
        best = 1e27
        stack = [(0, 0.0)]
        while stack:
            wi, entry = stack.pop()
            if policy == "near+cull" and entry > best:
                continue
            visits += 1
            inner = []
            for (blo, bhi, is_inner, payload, slot, a, sz) in nodes[wi]:
                t0, t1 = (blo - o) * inv, (bhi - o) * inv
                tn, tf = max(np.minimum(t0, t1).max(), tmin), min(np.maximum(t0, t1).min(), best)
                if tn > tf:
                    continue
                if is_inner:
                    inner.append((slot ^ octinv if policy == "octant" else -tn, payload, tn))
                else:
                    for prim in payload:
                        tests += 1
                        tt = tri_t(o, d, tmin, best, t[prim, 0], t[prim, 1], t[prim, 2])
                        if tt is not None:
                            best = tt
            for key, payload, tn in sorted(inner, key=lambda e: e[0]):
                stack.append((payload, tn))
        return visits, tests

    def trace_groups(o, d, tmin, cull):
        """The kernel's own stack discipline: the hit inner children of a node form one group in octant order; the
        front-most is entered at once, the rest is postponed as ONE stack entry.  cull = "none" (today), "group"
        (the entry also keeps the smallest entry distance of its members and is dropped whole once the best hit is
        nearer) or "child" (per-member distances: an upper bound for what distances on the stack can buy)."""
        inv = 1.0 / np.where(np.abs(d) > 1e-30, d, 1e-30)
        octinv = 7 ^ ((4 if d[0] < 0 else 0) | (2 if d[1] < 0 else 0) | (1 if d[2] < 0 else 0))
        visits = tests = 0
        best = 1e27
        stack = []
        group = [(0, 0.0)]
        while True:
            if not group:
                if not stack:
                    break
                group = stack.pop()
                if cull == "group" and min(e[1] for e in group) > best:
                    group = []
                    continue
            wi, entry = group.pop()          # front-most member
            if group:
                stack.append(group)
            group = []
            if cull == "child" and entry > best:
                continue
            visits += 1
            inner = []
            for (blo, bhi, is_inner, payload, slot, a, sz) in nodes[wi]:
                t0, t1 = (blo - o) * inv, (bhi - o) * inv
                tn, tf = max(np.minimum(t0, t1).max(), tmin), min(np.maximum(t0, t1).min(), best)
                if tn > tf:
                    continue
                if is_inner:
                    inner.append((slot ^ octinv, payload, tn))
                else:
                    for prim in payload:
                        tests += 1
                        tt = tri_t(o, d, tmin, best, t[prim, 0], t[prim, 1], t[prim, 2])
                        if tt is not None:
                            best = tt
            group = [(payload, tn) for key, payload, tn in sorted(inner, key=lambda e: e[0])]
        return visits, tests

    if "--closest" in sys.argv:
        # bounce-like rays: from a surface point into a random direction of the hemisphere
        crays = []
        for (o, d, tmin, tmax) in rays:
            v = rng.normal(size=3)
            v /= np.linalg.norm(v)
            crays.append((o, v if np.dot(v, d) > 0 else -v, tmin))
        for policy in ("octant", "near", "near+cull"):
            res = np.array([trace_closest(*r, policy) for r in crays], dtype=float)
            print(f"closest hit, {policy:9s} nodes/ray {res[:, 0].mean():5.2f}   prims/ray {res[:, 1].mean():5.2f}")
        for cull in ("none", "group", "child"):
            res = np.array([trace_groups(*r, cull) for r in crays], dtype=float)
            print(f"closest hit, kernel stack, cull {cull:6s} nodes/ray {res[:, 0].mean():5.2f}   prims/ray {res[:, 1].mean():5.2f}")
    print(f"{n} triangles, {len(nodes)} wide nodes, {len(rays)} shadow rays to {len(lights)} lights")
    for policy in ("octant", "near", "area", "overlap", "size"):
        res = np.array([trace(*r, policy) for r in rays], dtype=float)
        occ = res[:, 2] > 0
        print(f"{policy:8s} occluded {occ.mean() * 100:4.1f} %   nodes/ray {res[:, 0].mean():5.2f} (occluded {res[occ, 0].mean():5.2f}, "
              f"free {res[~occ, 0].mean():5.2f})   prims/ray {res[:, 1].mean():5.2f} (occluded {res[occ, 1].mean():5.2f})")


if __name__ == "__main__":
    main()
