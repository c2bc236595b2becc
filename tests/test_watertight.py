"""MOX_ACCEL_WATERTIGHT: the opt-in watertight ray-triangle test (Woop, Benthin, Wald 2013) that the north star
names.  It is not what the reference does (its mesh program calls the SDK's intersect_triangle, Geometry.cu:133,
which stays the default so that primitive ids are bit-exact), so it has its own oracle twin
(oracle/oracle.cpp::watertightTriangle) and its own properties:

  * no ray aimed at a vertex or an edge shared by triangles of a closed height field slips through (CPU + GPU);
  * GPU and oracle agree bit for bit on id / t / beta / gamma, on every hierarchy (GPU);
  * against the default test only epsilon ties change: |dt| <= 1e-5 t or an edge graze.
"""
import numpy as np
import pytest

from minimaloptix_b200 import structs as S


def height_field(n, seed):
    """n x n vertices over [0,1]^2 with a random height, 2 (n-1)^2 triangles sharing every interior edge.  Slopes stay
    below 0.2, so no ray of the tests below (at most ~40 degrees off the vertical) sees a silhouette: every ray crosses
    the field exactly once."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(0.0, 1.0, n, dtype=np.float32)
    x, y = np.meshgrid(xs, xs, indexing="ij")
    x = x + rng.uniform(-0.3, 0.3, size=x.shape).astype(np.float32) / np.float32(n)
    y = y + rng.uniform(-0.3, 0.3, size=y.shape).astype(np.float32) / np.float32(n)
    x[0, :], x[-1, :], y[:, 0], y[:, -1] = 0.0, 1.0, 0.0, 1.0
    z = (rng.uniform(-0.04, 0.04, size=x.shape) / n).astype(np.float32)
    v = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    idx = []
    for i in range(n - 1):
        for j in range(n - 1):
            a, b, c, d = i * n + j, (i + 1) * n + j, (i + 1) * n + j + 1, i * n + j + 1
            idx += [(a, b, c), (a, c, d)] if (i + j) % 2 else [(a, b, d), (b, c, d)]
    return v, np.asarray(idx, dtype=np.int32)


def rays_at_seams(v, tri, count, seed):
    """Rays from above the field aimed exactly at interior vertices and at points on shared edges."""
    rng = np.random.default_rng(seed)
    n = int(round(np.sqrt(len(v))))
    interior = np.asarray([i * n + j for i in range(2, n - 2) for j in range(2, n - 2)])
    tv = v[rng.choice(interior, size=count // 2)]
    t = tri[rng.integers(0, len(tri), size=count - count // 2)]
    a, b = v[t[:, 0]], v[t[:, 1]]
    keep = np.all((a[:, :2] > 0.1) & (a[:, :2] < 0.9) & (b[:, :2] > 0.1) & (b[:, :2] < 0.9), axis=1)
    s = rng.uniform(0, 1, size=(len(a), 1)).astype(np.float32)
    te = (a + s * (b - a)).astype(np.float32)[keep]
    target = np.concatenate([tv, te]).astype(np.float32)
    o = np.empty_like(target)
    o[:, 0:2] = rng.uniform(0.3, 0.7, size=(len(target), 2))
    o[:, 2] = rng.uniform(1.0, 2.0, size=len(target))
    d = (target - o).astype(np.float32)
    inv = (np.float32(1.0) / np.sqrt((d * d).sum(axis=1, dtype=np.float32))).astype(np.float32)
    d = (d * inv[:, None]).astype(np.float32)
    rays = np.empty((len(target), 8), dtype=np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = o, 1e-3, d, 1e27
    return rays


def rays_at_the_field(count, seed):
    """Rays from above aimed at random points of the field (generic position: seams are hit with probability 0)."""
    rng = np.random.default_rng(seed)
    target = np.zeros((count, 3), dtype=np.float32)
    target[:, 0:2] = rng.uniform(0.1, 0.9, size=(count, 2))
    o = np.empty_like(target)
    o[:, 0:2] = rng.uniform(0.3, 0.7, size=(count, 2))
    o[:, 2] = rng.uniform(1.0, 2.0, size=count)
    d = (target - o).astype(np.float32)
    inv = (np.float32(1.0) / np.sqrt((d * d).sum(axis=1, dtype=np.float32))).astype(np.float32)
    d = (d * inv[:, None]).astype(np.float32)
    rays = np.empty((count, 8), dtype=np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = o, 1e-3, d, 1e27
    return rays


def _field_context(make_ctx, flags, n=40, seed=3):
    v, tri = height_field(n, seed)
    ctx = make_ctx()
    ctx.set_globals(32, 32, 5)
    lam = S.LambertianParams()
    lam.albedo = S.float3(0.5, 0.5, 0.5)
    ctx.add_mesh(v, tri, S.MAT_LAMBERTIAN, lam)
    ctx.build_accel(flags)
    return ctx, v, tri


def test_oracle_watertight_closes_the_seams(orc):
    """Every ray aimed at a shared vertex / edge hits the field with the watertight test.  The SDK test (default)
    is allowed to leak — the count is printed; it is why the mode exists."""
    wt, v, tri = _field_context(orc.context, S.ACCEL_WATERTIGHT)
    sdk, _, _ = _field_context(orc.context, S.ACCEL_DEFAULT)
    rays = rays_at_seams(v, tri, 60000, 5)
    t_w, i_w, b_w, g_w = wt.trace_closest(rays)
    t_s, i_s, _, _ = sdk.trace_closest(rays)
    assert (i_w < 0).sum() == 0, "watertight test leaked"
    print("SDK intersect_triangle leaks on seams:", int((i_s < 0).sum()), "of", len(rays))
    # barycentrics are a partition of unity up to rounding, and t agrees with the SDK test where both hit
    assert np.all(b_w >= 0) and np.all(g_w >= 0) and np.all(b_w + g_w <= 1 + 1e-6)
    both = (i_s >= 0)
    assert both.sum() > 30000 and np.allclose(t_w[both], t_s[both], rtol=1e-5, atol=0)
    assert (i_s < 0).sum() > 0    # the deviation this mode exists for


def test_oracle_watertight_equals_sdk_off_the_seams(orc):
    """On generic rays the two tests pick the same triangle (ties aside) and agree on t / beta / gamma to rounding."""
    wt, v, tri = _field_context(orc.context, S.ACCEL_WATERTIGHT)
    sdk, _, _ = _field_context(orc.context, S.ACCEL_DEFAULT)
    rays = rays_at_the_field(50000, 9)
    t_w, i_w, b_w, g_w = wt.trace_closest(rays)
    t_s, i_s, b_s, g_s = sdk.trace_closest(rays)
    same = i_w == i_s
    assert same.mean() > 0.9999
    hit = same & (i_w >= 0)
    assert hit.sum() > 49000
    assert np.allclose(t_w[hit], t_s[hit], rtol=2e-5) and np.allclose(b_w[hit], b_s[hit], atol=2e-4) and np.allclose(g_w[hit], g_s[hit], atol=2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [S.ACCEL_DEFAULT, S.ACCEL_BINARY, S.ACCEL_LBVH], ids=["wide", "binary-ploc", "binary-lbvh"])
def test_gpu_watertight_matches_oracle_bit_for_bit(orc, gpu_backend, flags):
    o, v, tri = _field_context(lambda: orc.context(brute_force=True), S.ACCEL_WATERTIGHT, n=24)
    g, _, _ = _field_context(lambda: gpu_backend.context(0), S.ACCEL_WATERTIGHT | flags, n=24)
    seam = rays_at_seams(v, tri, 40000, 6)
    rays = np.concatenate([seam, rays_at_the_field(40000, 10)])
    to, io, bo, go = o.trace_closest(rays)
    tg, ig, bg, gg = g.trace_closest(rays)
    assert len(seam) > 30000 and (ig < 0).sum() == 0     # no leak on the GPU either (the first len(seam) rays aim at seams)
    assert np.array_equal(io, ig)
    for a, b in ((to, tg), (bo, bg), (go, gg)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.gpu
def test_gpu_watertight_render_matches_oracle(host, api_tables, orc, gpu_backend):
    """A Disney mesh scene rendered with the watertight test on both sides: equal ray counts, RMSE <= 1e-5; and the
    default test gives the same picture up to epsilon-tie paths."""
    sc = host.Scene.builtin("interior", 20000)
    imgs = {}
    for name, ctx, api, flags in (("orc", orc.context(), api_tables.oracle, S.ACCEL_WATERTIGHT),
                                  ("gpu", gpu_backend.context(0), api_tables.gpu, S.ACCEL_WATERTIGHT),
                                  ("gpu_sdk", gpu_backend.context(0), api_tables.gpu, S.ACCEL_DEFAULT)):
        sc.upload(api, ctx, 160, 90, 5)
        ctx.build_accel(flags)
        ctx.render(2, 17)
        st = ctx.stats()
        imgs[name] = (ctx.read_accum() / 2, st["rays_bounce"], st["rays_shadow"])
    assert imgs["orc"][1:] == imgs["gpu"][1:]
    rmse = float(np.sqrt(np.mean((imgs["orc"][0] - imgs["gpu"][0]) ** 2)))
    assert rmse <= 1e-5, rmse
    # watertight vs SDK test: the same paths except where a hit is an epsilon tie; the images differ in few pixels
    diff = np.abs(imgs["gpu"][0] - imgs["gpu_sdk"][0]).max(axis=-1)
    assert (diff > 1e-3).mean() < 0.01
