"""Known-answer tests that pin the oracle's restatement of the reference device code.

The reference ships no tests; these vectors were derived from its source text
(utils_device.h:8-34 RNG, Structures.h sizes, MinimalOptiX.cpp:49-52
quantisation) and are also stored in tests/golden/kat.json.
"""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from minimaloptix_b200 import structs as S

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


def test_tea16_vectors(orc):
    b = orc.backend()
    for v0, v1, want in GOLD["tea16"]:
        assert b.tea16(int(v0, 16), int(v1, 16)) == int(want, 16)


def test_lcg_sequence(orc):
    b = orc.backend()
    seed = C.c_int32(0)
    got = [b.lcg(C.byref(seed)) for _ in range(5)]
    assert got == [int(x, 16) for x in GOLD["lcg_from_0"]]
    assert seed.value & 0xFFFFFFFF == int(GOLD["lcg_final_seed"], 16)
    seed = C.c_int32(np.int32(np.uint32(0x80000001)).item())  # negative int seed
    got = [b.lcg(C.byref(seed)) for _ in range(3)]
    assert got == [int(x, 16) for x in GOLD["lcg_from_80000001"]]


def test_rand_from_tea00(orc):
    b = orc.backend()
    seed = C.c_int32(np.int32(np.uint32(b.tea16(0, 0))).item())
    got = [b.rand(C.byref(seed)) for _ in range(4)]
    assert got == pytest.approx(GOLD["rand_from_tea00"], rel=0, abs=1e-9)
    assert all(0.0 <= g < 1.0 for g in got)


def test_philox_known_answer(orc):
    # Random123 KAT for philox4x32-10: ctr=0,key=0 and the all-ones vector.
    b = orc.backend()
    out = (C.c_uint32 * 4)()
    b.philox((C.c_uint32 * 4)(0, 0, 0, 0), (C.c_uint32 * 2)(0, 0), out)
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    b.philox((C.c_uint32 * 4)(*[0xFFFFFFFF] * 4), (C.c_uint32 * 2)(0xFFFFFFFF, 0xFFFFFFFF), out)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_struct_sizes():
    for t, size in S.SIZES.items():
        assert C.sizeof(t) == size
    assert S.QuadParams.v1.offset == 16 and S.QuadParams.v2.offset == 28 and S.QuadParams.anchor.offset == 40
    assert S.DisneyParams.color.offset == 4 and S.DisneyParams.emission.offset == 16
    assert S.DisneyParams.metallic.offset == 28 and S.DisneyParams.roughness.offset == 40
    assert S.DisneyParams.brdfType.offset == 68
    assert S.LightParams.u.offset == 36 and S.LightParams.area.offset == 60 and S.LightParams.shape.offset == 68


def _f3(*v):
    return (C.c_float * 3)(*v)


def test_refract_snell_and_tir(orc):
    b = orc.backend()
    out = _f3(0, 0, 0)
    # 45 degree incidence into ior 1.5: sin(t) = sin(45)/1.5
    i = _f3(math.sqrt(0.5), -math.sqrt(0.5), 0)
    assert b.refract(i, _f3(0, 1, 0), 1.5, out) == 1
    assert abs(out[0] - math.sqrt(0.5) / 1.5) < 1e-6 and out[1] < 0
    assert abs(out[0] ** 2 + out[1] ** 2 + out[2] ** 2 - 1) < 1e-6
    # leaving glass at a grazing angle: total internal reflection.  The caller flips the
    # normal towards the ray and passes 1/ior (Material.cu:83-87).
    i = _f3(0.9, math.sqrt(1 - 0.81), 0)
    assert b.refract(i, _f3(0, -1, 0), 1 / 1.5, out) == 0
    assert list(out) == [0, 0, 0]


def test_fresnel_normal_incidence(orc):
    b = orc.backend()
    r = b.fresnel(1.0, 1.0, 1.5)
    assert abs(r - ((1.5 - 1) / (1.5 + 1)) ** 2) < 1e-6


def test_refine_hitpoint_sides(orc):
    b = orc.backend()
    back, front = _f3(0, 0, 0), _f3(0, 0, 0)
    # ray going down onto the plane y = 1 with normal +y: front is above, back below
    b.refine_hitpoint(_f3(0.3, 1.00001, 0.2), _f3(0, -1, 0), _f3(0, 1, 0), _f3(5, 1, 5), back, front)
    assert front[1] > 1.0 > back[1]
    assert abs(front[1] - 1.0) < 2e-3 and abs(back[1] - 1.0) < 2e-3
    # x and z are far from 0 and the normal has no x/z part: unchanged
    assert front[0] == pytest.approx(0.3, abs=1e-6) and front[2] == pytest.approx(0.2, abs=1e-6)
    # coordinates near zero use the absolute epsilon branch
    b.refine_hitpoint(_f3(0.0, 0.0, 0.0), _f3(0, -1, 0), _f3(0, 1, 0), _f3(0, 0, 0), back, front)
    assert front[1] == pytest.approx(1e-4, rel=1e-3) and back[1] == pytest.approx(-1e-4, rel=1e-3)


def test_disney_pdf_integrates_to_one(orc):
    """disneyPdf is a density over directions L.  Over the FULL sphere of L the GTR2 half-vector
    lobe integrates to 1 and the |N.L|/pi diffuse lobe to 2 (the abs() mirrors it below the
    surface), so the integral is specularRatio + 2 * diffuseRatio = 1 + diffuseRatio."""
    b = orc.backend()
    d = S.DisneyParams()
    d.color = S.float3(0.8, 0.6, 0.4)
    d.specular, d.roughness, d.sheenTint, d.clearcoatGloss, d.metallic = 0.5, 0.5, 0.5, 1.0, 0.3
    N = np.array([0.0, 0.0, 1.0])
    V = np.array([math.sin(0.6), 0.0, math.cos(0.6)])
    rng = np.random.default_rng(1)
    n = 200000
    L = rng.normal(size=(n, 3))
    L /= np.linalg.norm(L, axis=1, keepdims=True)
    total = 0.0
    for k in range(n):
        l = L[k]
        h = l + V
        h /= np.linalg.norm(h)
        total += b.disney_pdf(C.byref(d), _f3(*N), _f3(*l), _f3(*V), _f3(*h))
    integral = total / n * 4 * math.pi
    diffuse_ratio = 0.5 * (1 - 0.3)
    assert integral == pytest.approx(1 + diffuse_ratio, rel=0.03)


def test_disney_eval_reciprocity_and_positivity(orc):
    b = orc.backend()
    d = S.DisneyParams()
    d.color = S.float3(0.8, 0.6, 0.4)
    d.specular, d.roughness, d.sheenTint, d.clearcoatGloss = 0.5, 0.4, 0.5, 1.0
    N = _f3(0, 0, 1)
    L = np.array([0.3, 0.2, 0.9]); L /= np.linalg.norm(L)
    V = np.array([-0.5, 0.1, 0.8]); V /= np.linalg.norm(V)
    H = L + V; H /= np.linalg.norm(H)
    o1, o2 = _f3(0, 0, 0), _f3(0, 0, 0)
    bc = _f3(0.8, 0.6, 0.4)
    b.disney_eval(C.byref(d), bc, N, _f3(*L), _f3(*V), _f3(*H), o1)
    b.disney_eval(C.byref(d), bc, N, _f3(*V), _f3(*L), _f3(*H), o2)
    assert all(x > 0 for x in o1)
    assert list(o1) == pytest.approx(list(o2), rel=1e-5)  # isotropic settings: f(L,V) = f(V,L)


def test_disney_sample_draw_order(orc):
    """metallic = 1 -> diffuseRatio 0 -> specular lobe: exactly 3 draws (lobe, phi, xi)."""
    b = orc.backend()
    d = S.DisneyParams()
    d.color = S.float3(1, 1, 1)
    d.metallic, d.roughness, d.clearcoatGloss = 1.0, 0.3, 1.0
    seed = C.c_int32(12345)
    L, H = _f3(0, 0, 0), _f3(0, 0, 0)
    b.disney_sample(C.byref(seed), C.byref(d), _f3(0, 0, 1), _f3(0, 0.6, 0.8), L, H)
    ref = C.c_int32(12345)
    for _ in range(3):
        b.lcg(C.byref(ref))
    assert seed.value == ref.value
    assert abs(sum(x * x for x in L) - 1) < 1e-5 and abs(sum(x * x for x in H) - 1) < 1e-5
