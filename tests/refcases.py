"""TEST INFRASTRUCTURE: the cases that pin oracle/ to the REFERENCE'S OWN device code.

`oracle/_ref/libref_render.so` is the reference's Camera.cu / Geometry.cu / Material.cu / miss.cu /
disney.h / utils_device.h compiled unchanged with g++ behind an OptiX shim (oracle/ref_shim/,
`make ref`; exists only where /root/reference does).  This module describes inputs once and
evaluates them on any library exporting the render ABI + the helper hooks under a prefix
(`ref_` = compiled reference, `orc_` = the oracle):

  * helper_vectors(): tea/lcg/rand, randInUnitSphere/Disk, folkPayload seed, fresnel, offset,
    refineHitpoint, GTR1/GTR2/GTR2Aniso/schlickFresnel/smithGGgx/smithGGgxAniso/powerHeuristic/
    srgb2lin, disneySample/Pdf/Eval
  * scene_cases(): per-program results — closest hits with all five attributes, bounding-box
    programs, shadow transmittance, and whole images through camera() -> programs -> accumulate
    (rows a-4..a-16)

scripts/make_render_golden.py stores the `ref_` results in tests/golden/render_ref.npz;
tests/test_ref_render.py compares `orc_` with them bit for bit.

Evaluation order.  g++ gives the first draw of `light.u * rand(s) + light.v * rand(s)`
(Material.cu:180) to v, nvcc to u; the oracle follows nvcc by default and g++ with
orc_set_quad_light_draw_order(1).  Every other draw inside an argument list is pinned left to
right by the shim (braced initialisation).  The goldens are therefore generated, and compared,
with draw order 1; the default order differs from it by that swap alone.
"""
import ctypes as C
import os

import numpy as np

from minimaloptix_b200 import structs as S
from minimaloptix_b200._binding import Backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_render.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "render_ref.npz")

_vp, _u32, _i32, _f = C.c_void_p, C.c_uint32, C.c_int32, C.c_float
_i32p, _f3 = C.POINTER(C.c_int32), C.c_float * 3
_DP = C.POINTER(S.DisneyParams)

HOOKS = {
    "tea16": (_u32, [_u32, _u32]),
    "lcg": (_u32, [_i32p]),
    "rand": (_f, [_i32p]),
    "rand_in_unit_sphere": (None, [_i32p, _f3]),
    "rand_in_unit_disk": (None, [_i32p, _f3]),
    "fork_seed": (_i32, [_i32, _i32]),
    "fresnel": (_f, [_f, _f, _f]),
    "offset": (None, [_f3, _f3, _f3]),
    "refine_hitpoint": (None, [_f3, _f3, _f3, _f3, _f3, _f3]),
    "gtr1": (_f, [_f, _f]),
    "gtr2": (_f, [_f, _f]),
    "gtr2_aniso": (_f, [_f, _f, _f, _f, _f]),
    "schlick_fresnel": (_f, [_f]),
    "smith_ggx": (_f, [_f, _f]),
    "smith_ggx_aniso": (_f, [_f, _f, _f, _f, _f]),
    "power_heuristic": (_f, [_f, _f]),
    "srgb2lin": (None, [_f3, _f3]),
    "disney_eval": (None, [_DP, _f3, _f3, _f3, _f3, _f3, _f3]),
    "disney_pdf": (_f, [_DP, _f3, _f3, _f3, _f3]),
    "disney_sample": (None, [_i32p, _DP, _f3, _f3, _f3, _f3]),
    "refract": (C.c_int, [_f3, _f3, _f, _f3]),
    "trace_closest_attrs": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp]),
    "set_threads": (C.c_int, [_vp, C.c_int]),
}
ORC_ONLY = {"prim_bounds": (C.c_int, [_vp, _u32, C.c_float * 6, C.POINTER(C.c_int)]),
            "set_brute_force": (C.c_int, [_vp, C.c_int]),
            "set_quad_light_draw_order": (C.c_int, [_vp, C.c_int])}
REF_ONLY = {"prim_bounds": (C.c_int, [_vp, _u32, C.c_float * 6]),
            "exception": (C.c_int, [_vp, _u32, _u32])}


def have_ref():
    return os.path.exists(REF_LIB)


def load(which):
    """A Backend with the helper hooks bound: which = 'ref' (compiled reference) or 'orc' (oracle)."""
    if which == "ref":
        extra = dict(HOOKS); extra.update(REF_ONLY)
        return Backend(REF_LIB, "ref_", extra=extra)
    import oracle
    extra = dict(HOOKS); extra.update(ORC_ONLY)
    b = Backend(oracle.ORACLE_LIB, "orc_", extra=extra)
    return b


def new_context(b, threads=0):
    ctx = b.context(0)
    b.set_threads(ctx.h, threads)
    if b.prefix == "orc_":
        b.set_brute_force(ctx.h, 1)            # primitive-id order, like the harness
        b.set_quad_light_draw_order(ctx.h, 1)  # g++'s order of Material.cu:180 (see module docstring)
    return ctx


# ------------------------------------------------------------------------------------------ helpers
def _unit(rng, n):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v = v.astype(np.float32)
    inv = (np.float32(1) / np.sqrt((v * v).sum(axis=1, dtype=np.float32))).astype(np.float32)
    return (v * inv[:, None]).astype(np.float32)


def random_disney(rng):
    d = S.DisneyParams()
    d.albedoID = 0
    d.color = S.float3(*rng.uniform(0.02, 1, 3))
    d.emission = S.float3(0, 0, 0)
    for k in ("metallic", "subsurface", "specular", "roughness", "specularTint", "anisotropic", "sheen", "sheenTint",
              "clearcoat", "clearcoatGloss"):
        setattr(d, k, float(rng.uniform(0, 1)))
    if rng.uniform() < 0.2:
        d.roughness = float(rng.choice([0.0, 0.001, 1.0]))
    if rng.uniform() < 0.2:
        d.clearcoatGloss = float(rng.choice([0.0, 1.0]))
    d.brdfType = S.NORMAL
    return d


def helper_vectors(b, n=256, seed=2024):
    """name -> array of results of every device helper on seeded random inputs."""
    rng = np.random.default_rng(seed)
    out = {}
    a = rng.integers(0, 1 << 32, size=(n, 2), dtype=np.uint64).astype(np.uint32)
    a[:8] = [[0, 0], [1, 0], [0, 1], [1, 1], [12345, 67890], [1037760, 0x12345678], [0xFFFFFFFF, 0xFFFFFFFF], [131328, 1]]
    out["tea16"] = np.array([b.tea16(int(x), int(y)) for x, y in a], dtype=np.uint32)
    seeds = rng.integers(-(1 << 31), 1 << 31, size=n, dtype=np.int64).astype(np.int32)
    seeds[:3] = [0, -2147483647, 0x741c187d]
    lcg, rnd, sph, dsk, sd = [], [], [], [], []
    for s0 in seeds:
        s = C.c_int32(int(s0))
        lcg.append([b.lcg(C.byref(s)) for _ in range(4)] + [s.value & 0xFFFFFFFF])
        s = C.c_int32(int(s0))
        rnd.append([b.rand(C.byref(s)) for _ in range(4)])
        s = C.c_int32(int(s0)); o = _f3()
        b.rand_in_unit_sphere(C.byref(s), o); sph.append(list(o) + [np.float32(0)]); sd.append(s.value)
        s = C.c_int32(int(s0)); o = _f3()
        b.rand_in_unit_disk(C.byref(s), o); dsk.append(list(o)); sd.append(s.value)
    out["lcg"] = np.array(lcg, dtype=np.uint32)
    out["rand"] = np.array(rnd, dtype=np.float32)
    out["rand_in_unit_sphere"] = np.array(sph, dtype=np.float32)
    out["rand_in_unit_disk"] = np.array(dsk, dtype=np.float32)
    out["rand_in_unit_seeds_after"] = np.array(sd, dtype=np.int64).astype(np.int32)
    out["fork_seed"] = np.array([b.fork_seed(int(s0), int(d)) for s0, d in zip(seeds, rng.integers(1, 300, n))], dtype=np.int32)

    ci, ct = rng.uniform(0, 1, n).astype(np.float32), rng.uniform(0, 1, n).astype(np.float32)
    ior = rng.choice([1.5, 1 / 1.5, 1.45, 1 / 1.45, 2.4], n).astype(np.float32)
    out["fresnel"] = np.array([b.fresnel(float(x), float(y), float(z)) for x, y, z in zip(ci, ct, ior)], dtype=np.float32)

    # hit-point refinement: points of very different magnitude (both branches of offset())
    scale = (10.0 ** rng.uniform(-6, 3, size=(n, 1))).astype(np.float32)
    hit = (rng.uniform(-1, 1, size=(n, 3)).astype(np.float32) * scale).astype(np.float32)
    hit[:4] = [[0, 0, 0], [1e-5, -1e-5, 5e-5], [1e-4, 1e-4, -1e-4], [123.5, -0.25, 1e-3]]
    nrm, dirs = _unit(rng, n), _unit(rng, n)
    p0 = (hit + rng.normal(size=(n, 3)).astype(np.float32) * scale * np.float32(0.1)).astype(np.float32)
    off, ref = [], []
    for k in range(n):
        o = _f3(); b.offset(_f3(*hit[k]), _f3(*nrm[k]), o); off.append(list(o))
        bk, fr = _f3(), _f3()
        b.refine_hitpoint(_f3(*hit[k]), _f3(*dirs[k]), _f3(*nrm[k]), _f3(*p0[k]), bk, fr); ref.append(list(bk) + list(fr))
    out["offset"] = np.array(off, dtype=np.float32)
    out["refine_hitpoint"] = np.array(ref, dtype=np.float32)

    u = rng.uniform(-0.2, 1.2, n).astype(np.float32)
    al = np.concatenate([rng.uniform(0.001, 1.2, n - 4), [1.0, 0.001, 0.1, 1.5]]).astype(np.float32)
    c5 = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    ax, ay = rng.uniform(0.001, 1, n).astype(np.float32), rng.uniform(0.001, 1, n).astype(np.float32)
    out["gtr1"] = np.array([b.gtr1(float(x), float(y)) for x, y in zip(ci, al)], dtype=np.float32)
    out["gtr2"] = np.array([b.gtr2(float(x), float(y)) for x, y in zip(ci, al)], dtype=np.float32)
    out["gtr2_aniso"] = np.array([b.gtr2_aniso(float(c[0]), float(c[1]), float(c[2]), float(x), float(y)) for c, x, y in zip(c5, ax, ay)], dtype=np.float32)
    out["schlick_fresnel"] = np.array([b.schlick_fresnel(float(x)) for x in u], dtype=np.float32)
    out["smith_ggx"] = np.array([b.smith_ggx(float(x), float(y)) for x, y in zip(ci, al)], dtype=np.float32)
    out["smith_ggx_aniso"] = np.array([b.smith_ggx_aniso(float(c[0]), float(c[1]), float(c[2]), float(x), float(y)) for c, x, y in zip(c5, ax, ay)], dtype=np.float32)
    out["power_heuristic"] = np.array([b.power_heuristic(float(x), float(y)) for x, y in zip(ci * 50, ct * 50)], dtype=np.float32)
    lin = []
    for k in range(n):
        o = _f3(); b.srgb2lin(_f3(*rng.uniform(0, 1, 3)), o); lin.append(list(o))
    out["srgb2lin"] = np.array(lin, dtype=np.float32)

    # Disney: N random, V in N's hemisphere; sample L/H with the reference's sampler, then pdf and
    # eval for that pair and for an independent L in the hemisphere (the NEE use).
    N = _unit(rng, n)
    V = _unit(rng, n)
    flip = (N * V).sum(axis=1) < 0
    V[flip] = -V[flip]
    L2 = _unit(rng, n)
    flip = (N * L2).sum(axis=1) < 0
    L2[flip] = -L2[flip]
    smp, pdf, ev, rfr = [], [], [], []
    for k in range(n):
        d = random_disney(rng)
        base = _f3(*rng.uniform(0, 1, 3))
        s = C.c_int32(int(seeds[k])); L, H = _f3(), _f3()
        b.disney_sample(C.byref(s), C.byref(d), _f3(*N[k]), _f3(*V[k]), L, H)
        smp.append(list(L) + list(H))
        pdf.append(b.disney_pdf(C.byref(d), _f3(*N[k]), L, _f3(*V[k]), H))
        o = _f3(); b.disney_eval(C.byref(d), base, _f3(*N[k]), L, _f3(*V[k]), H, o); e1 = list(o)
        h2 = L2[k] + V[k]
        h2 = (h2 / np.linalg.norm(h2)).astype(np.float32)
        pdf.append(b.disney_pdf(C.byref(d), _f3(*N[k]), _f3(*L2[k]), _f3(*V[k]), _f3(*h2)))
        o = _f3(); b.disney_eval(C.byref(d), base, _f3(*N[k]), _f3(*L2[k]), _f3(*V[k]), _f3(*h2), o)
        ev.append(e1 + list(o))
        o = _f3(); ok = b.refract(_f3(*(-V[k])), _f3(*N[k]), float(ior[k]), o); rfr.append(list(o) + [np.float32(ok)])
    out["disney_sample"] = np.array(smp, dtype=np.float32)
    out["disney_pdf"] = np.array(pdf, dtype=np.float32)
    out["disney_eval"] = np.array(ev, dtype=np.float32)
    out["sdk_refract"] = np.array(rfr, dtype=np.float32)
    return out


# ------------------------------------------------------------------------------------------ scenes
def _uv_sphere(center, radius, nu=10, nv=6):
    """Tessellated sphere with smooth normals and uvs (faces v/vt/vn like an OBJ)."""
    v, n, uv, f = [], [], [], []
    for j in range(nv + 1):
        th = np.pi * j / nv
        for i in range(nu + 1):
            ph = 2 * np.pi * i / nu
            d = np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)])
            v.append(np.array(center) + radius * d); n.append(d); uv.append([i / nu * 3.0, j / nv * 2.0])
    for j in range(nv):
        for i in range(nu):
            a, b_ = j * (nu + 1) + i, j * (nu + 1) + i + 1
            c, d = (j + 1) * (nu + 1) + i, (j + 1) * (nu + 1) + i + 1
            if j > 0:
                f.append([a, b_, c])
            if j < nv - 1:
                f.append([b_, d, c])
    return (np.array(v, np.float32), np.array(n, np.float32), np.array(uv, np.float32), np.array(f, np.int32))


def _disney(**kw):
    d = S.DisneyParams()
    d.albedoID = 0
    d.color = S.float3(1, 1, 1); d.emission = S.float3(0, 0, 0)
    d.specular, d.roughness, d.sheenTint, d.clearcoatGloss = 0.5, 0.5, 0.5, 1.0   # initDisneyParams, utils_host.cpp:101-116
    d.brdfType = S.NORMAL
    for k, val in kw.items():
        setattr(d, k, S.float3(*val) if k in ("color", "emission") else val)
    return d


def build_zoo(ctx, host, w, h, depth=5, textured=True):
    """Every program in one scene: Disney NORMAL with all lobes switched on (textured mesh with
    normals + uvs, a metallic anisotropic tessellated sphere, a clearcoat / sheen / subsurface /
    emissive analytic sphere), one Disney GLASS sphere, lambertian / metal / glass analytic
    primitives, a SPHERE light and a QUAD light.

    Primitive order is chosen so that the reference's order-dependent shadow any-hit
     and the product's order-independent rule give the same answer: Disney NORMAL
    primitives first, then the single Disney GLASS sphere (one accepted hit per ray), then the
    primitives that have no any-hit program."""
    ctx.set_globals(w, h, depth, bg=(0.3, 0.4, 0.5))
    ctx.set_camera(host.set_cam_params((0.3, 2.2, 6.0), (0, 0.8, 0), (0, 1, 0), 38, w / h, 0.08, 6.0))
    rng = np.random.default_rng(7)
    tex = np.ones((8, 8, 4), np.float32)
    tex[..., :3] = rng.uniform(0.05, 1, size=(8, 8, 3))
    tid = ctx.add_texture(tex)
    # floor: two triangles, uvs beyond [0,1] (REPEAT), normals tilted per vertex
    fv = np.array([[-4, 0, -4], [4, 0, -4], [4, 0, 4], [-4, 0, 4]], np.float32)
    fn = np.array([[0.05, 1, 0], [0, 1, 0.05], [-0.05, 1, 0], [0, 1, -0.05]], np.float32)
    fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    fuv = np.array([[0, 0], [2.5, 0], [2.5, 2.5], [0, 2.5]], np.float32)
    fi = np.array([[0, 2, 1], [0, 3, 2]], np.int32)
    floor = _disney(albedoID=tid, roughness=0.6, sheen=0.3) if textured else _disney(color=(0.6, 0.55, 0.5), roughness=0.6, sheen=0.3)
    ctx.add_mesh(fv, fi, S.MAT_DISNEY, floor, normals=fn, n_idx=fi, texcoords=fuv, t_idx=fi)
    v, n, uv, f = _uv_sphere((-1.6, 0.8, 0.2), 0.8)
    ctx.add_mesh(v, f, S.MAT_DISNEY, _disney(color=(0.9, 0.6, 0.2), metallic=0.8, roughness=0.3, anisotropic=0.6, specularTint=0.4),
                 normals=n, n_idx=f, texcoords=uv, t_idx=f)
    # a mesh without normals / uvs (flat shading path of meshIntersect), one degenerate face
    bv = np.array([[1.2, 0, -2.2], [2.8, 0, -2.2], [2.0, 1.8, -2.0], [2.0, 0, -1.0], [2.0, 0, -1.0]], np.float32)
    bi = np.array([[0, 1, 2], [1, 3, 2], [3, 0, 2], [3, 4, 4]], np.int32)
    ctx.add_mesh(bv, bi, S.MAT_DISNEY, _disney(color=(0.2, 0.7, 0.3), subsurface=0.5, roughness=0.8))
    ctx.add_sphere(S.SphereParams(0.6, S.float3(0.2, 0.6, 1.2), S.float3()), S.MAT_DISNEY,
                   _disney(color=(0.7, 0.2, 0.2), clearcoat=1.0, clearcoatGloss=0.5, sheen=0.6, subsurface=0.4, specularTint=0.5,
                           emission=(0.05, 0.0, 0.02), roughness=0.35))
    ctx.add_sphere(S.SphereParams(0.55, S.float3(1.6, 0.55, 0.4), S.float3()), S.MAT_DISNEY,
                   _disney(color=(0.9, 0.8, 0.7), brdfType=S.GLASS))
    ctx.add_sphere(S.SphereParams(0.4, S.float3(-0.4, 0.4, 2.4), S.float3()), S.MAT_LAMBERTIAN, S.LambertianParams(S.float3(0.2, 0.3, 0.8)))
    ctx.add_sphere(S.SphereParams(0.4, S.float3(0.9, 0.4, 2.6), S.float3()), S.MAT_METAL, S.MetalParams(S.float3(0.8, 0.7, 0.5), 0.25))
    ctx.add_sphere(S.SphereParams(0.35, S.float3(-1.5, 0.35, 2.2), S.float3()), S.MAT_GLASS, S.GlassParams(S.float3(0.95, 1.0, 0.95), 1.5))
    ctx.add_quad(host.set_quad_params((-4, 0, -4), (8, 0, 0), (0, 4, 0)), S.MAT_LAMBERTIAN, S.LambertianParams(S.float3(0.7, 0.7, 0.6)))
    lights = []
    ls = S.LightParams()
    ls.position, ls.emission, ls.radius, ls.shape = S.float3(-2.5, 3.2, 1.5), S.float3(9, 8, 7), 0.35, S.SPHERE
    ls.area = float(np.float32(4.0) * np.float32(np.pi) * np.float32(0.35) * np.float32(0.35))   # scene.cpp:87
    ctx.add_sphere(S.SphereParams(0.35, ls.position, S.float3()), S.MAT_LIGHT, ls)
    lights.append(ls)
    lq = S.LightParams()
    pos, u, vv = np.array([1.0, 3.5, -0.5], np.float32), np.array([1.2, 0, 0], np.float32), np.array([0, 0, 1.0], np.float32)
    nrm = np.cross(u, vv)
    lq.position, lq.u, lq.v = S.float3(*pos), S.float3(*u), S.float3(*vv)
    lq.area = float(np.linalg.norm(nrm)); lq.normal = S.float3(*(nrm / np.linalg.norm(nrm)))  # scene.cpp:74-82: faces down
    lq.emission, lq.shape = S.float3(6, 6, 6), S.QUAD
    ctx.add_quad(host.set_quad_params(tuple(pos), tuple(u), tuple(vv)), S.MAT_LIGHT, lq)
    lights.append(lq)
    ctx.set_lights(lights)
    ctx.build_accel()
    return dict(bbox_lo=(-4, 0, -4), bbox_hi=(4, 4, 4))


def build_builtin(name):
    def f(ctx, host, w, h, depth=5, tables={}):
        sc = host.Scene.builtin(name)
        lib = ctx.b.path
        if lib not in tables:
            tables[lib] = host.ApiTable(lib, ctx.b.prefix)
        sc.upload(tables[lib], ctx, w, h, depth)
        ctx.build_accel()
        return dict(bbox_lo=(-12, -1, -12), bbox_hi=(12, 8, 12))
    return f


def build_cornell(ctx, host, w, h, depth=5, tables={}):
    sc = host.Scene.load(os.path.join(ROOT, "scenes", "cornell"), "cornell")
    lib = ctx.b.path
    if lib not in tables:
        tables[lib] = host.ApiTable(lib, ctx.b.prefix)
    sc.upload(tables[lib], ctx, w, h, depth)
    ctx.build_accel()
    return dict(bbox_lo=(0.05, 0.05, 0.05), bbox_hi=(5.4, 5.4, 5.5))


# name -> (builder, width, height, spp, depth, seed, rays for the closest-hit / shadow queries)
SCENES = {
    "spheres_lens": (build_builtin("spheres_lens"), 64, 36, 3, 5, 0xC0FFEE, 1500),
    "spheres_pinhole": (build_builtin("spheres_pinhole"), 64, 36, 2, 256, 0xC0FFEE, 0),
    "random_spheres": (build_builtin("random_spheres"), 64, 36, 2, 5, 0x5EED, 1500),
    "cornell": (build_cornell, 64, 64, 4, 5, 0xC0FFEE, 3000),
    "zoo": (build_zoo, 72, 48, 4, 5, 0xD15EA5E, 4000),
    "zoo_depth1": (build_zoo, 36, 24, 2, 1, 11, 0),
}


def _rays(n, lo, hi, seed, tmax):
    rng = np.random.default_rng(seed)
    r = np.empty((n, 8), np.float32)
    r[:, 0:3] = rng.uniform(lo, hi, size=(n, 3))
    r[:, 4:7] = _unit(rng, n)
    r[:, 3], r[:, 7] = 1e-3, tmax
    return r


def scene_case(b, host, name, threads=0, size=None, spp=None):
    """Evaluate one scene on backend `b`: image after `spp` launches, ray counts, closest hits +
    attributes, shadow transmittance, bounding boxes."""
    build, w, h, spp0, depth, seed, nrays = SCENES[name]
    if size:
        w, h = size
    spp = spp or spp0
    ctx = new_context(b, threads)
    info = build(ctx, host, w, h, depth)
    ctx.render(spp, seed)
    st = ctx.stats()
    out = {"image": ctx.read_accum(),
           "ray_counts": np.array([st["rays_primary"], st["rays_bounce"], st["rays_shadow"]], dtype=np.int64)}
    if nrays:
        rays = _rays(nrays, info["bbox_lo"], info["bbox_hi"], 99, 1e27)
        hits = np.zeros((nrays, 4), np.float32)
        attrs = np.zeros((nrays, 15), np.float32)
        rc = b.trace_closest_attrs(ctx.h, rays.ctypes.data_as(_vp), nrays, hits.ctypes.data_as(_vp), attrs.ctypes.data_as(_vp))
        assert rc == 0
        out["hits"], out["attrs"] = hits, attrs
        out["shadow"] = ctx.trace_shadow(_rays(nrays, info["bbox_lo"], info["bbox_hi"], 98, 2.5))
        nprim = st["n_prims"]
        bounds = np.zeros((nprim, 7), np.float32)
        for p in range(nprim):
            o = (C.c_float * 6)()
            if b.prefix == "orc_":
                v = C.c_int()
                b.prim_bounds(ctx.h, p, o, C.byref(v))
                bounds[p] = list(o) + [v.value]
            else:
                b.prim_bounds(ctx.h, p, o)
                bounds[p] = list(o) + [1.0 if (o[0] <= o[3] and o[1] <= o[4] and o[2] <= o[5]) else 0.0]
        out["bounds"] = bounds
    ctx.close()
    return out


def all_cases(b, host, threads=0):
    res = {"helpers/" + k: v for k, v in helper_vectors(b).items()}
    for name in SCENES:
        for k, v in scene_case(b, host, name, threads).items():
            res[f"{name}/{k}"] = v
    return res


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return np.array_equal(a.view(np.uint8), b.view(np.uint8))


def bits_equal_rows(a, b):
    """Per-row bit equality of two float32 arrays of the same shape."""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)).reshape(len(a), -1).all(axis=1)
