"""The numpy model of the wide-BVH collapse (scripts/collapse_study.py, scripts/shadow_order_study.py) — the CPU
statement of what bvh_wide.cu::k_collapse_dp / k_collapse_level compute on the GPU: the dynamic programme's cost is the
cost of the collapse its decisions produce, that collapse is a valid 8-wide BVH, and it is never worse than the greedy
one.  (The GPU kernels themselves are covered by the parity tests: any valid tree gives the same hits.)"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "scripts"))


@pytest.mark.parametrize("n,seed,radius", [(2, 1, 4), (3, 2, 4), (17, 3, 4), (200, 4, 8), (1500, 5, 16)])
def test_optimal_collapse_is_valid_and_matches_its_cost(n, seed, radius):
    import collapse_study as cs
    import shadow_order_study as so
    rng = np.random.default_rng(seed)
    c = rng.random((n, 3))
    half = rng.random((n, 3)) * 0.05
    lo, hi = c - half, c + half
    order = np.argsort(cs.morton(c), kind="stable")
    lo, hi = lo[order], hi[order]
    L, R, nlo, nhi, size = cs.ploc(lo, hi, radius)
    root = len(L) - 1
    assert len(L) == 2 * n - 1 and size[root] == n
    assert all(L[k] < k and R[k] < k for k in range(n, len(L)))       # children are created before their parent
    A = cs.area(nlo, nhi) / cs.area(nlo[root], nhi[root])
    kbest, inherit = so.decisions(L, R, A, size, n)
    wide, index_of = so.collapse(L, R, size, n, root, kbest, inherit)
    # a valid 8-wide BVH: 2..8 children per node, leaf children of <= 2 primitives, every primitive exactly once
    seen, cost = [], 0.0
    binary_of = {v: k for k, v in index_of.items()}
    for wi, ch in enumerate(wide):
        assert 2 <= len(ch) <= cs.WIDTH
        cost += cs.C_NODE * A[binary_of[wi]]
        for node in ch:
            if size[node] > cs.LEAF_MAX:
                assert node in index_of
            else:
                prims = so.leaves_of(L, R, n, node)
                assert 1 <= len(prims) <= cs.LEAF_MAX
                seen += prims
                cost += cs.C_PRIM * A[node] * len(prims)
    assert sorted(seen) == list(range(n))
    assert len(wide) == len(index_of)
    # the programme's optimum is the cost of that tree (the root is forced to be a wide node), and beats greedy
    c_root_internal = cost
    assert np.isclose(c_root_internal, cs.optimal_cost(L, R, A, size, root, n), rtol=1e-9) or size[root] <= cs.LEAF_MAX
    greedy, _, _ = cs.greedy_cost(L, R, A, size, root)
    assert c_root_internal <= greedy * (1 + 1e-9)
