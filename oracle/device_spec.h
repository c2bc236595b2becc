// oracle/device_spec.h — TEST INFRASTRUCTURE (CPU oracle), never linked into the product.
//
// Scalar restatement of the reference's device helpers:
//   RNG                      utils_device.h:8-52   (tea<N>, lcg, rand, randInUnitSphere/Disk)
//   fresnel                  utils_device.h:63-67
//   hit-point refinement     utils_device.h:72-128 (intersectPlane, offset, refineHitpoint)
//   microfacet helpers       utils_device.h:130-185
//   Disney sample/pdf/eval   disney.h:9-91
// Where C++ leaves the evaluation order of rand() calls unspecified the
// order is pinned left-to-right with named temporaries; the CUDA side pins the same order.
// A second generator (Philox4x32-10) is the north-star "per-pixel Philox" mode; it is
// restated here so both sides can be compared path-for-path in that mode too.
#pragma once
#include "vecmath.h"
#include "../include/mox_structs.h"

namespace orc {

// ---------------------------------------------------------------- RNG
template <unsigned int N>
inline unsigned int tea(unsigned int val0, unsigned int val1) {
  unsigned int v0 = val0, v1 = val1, s0 = 0;
  for (unsigned int n = 0; n < N; n++) {
    s0 += 0x9e3779b9;
    v0 += ((v1 << 4) + 0xa341316c) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4);
    v1 += ((v0 << 4) + 0xad90777d) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761e);
  }
  return v0;
}

// Philox4x32-10 (Salmon et al. 2011), one block.
inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Per-path generator state.  mode 0 (REF): `seed` is Payload.randSeed, advanced by lcg().
// mode 1 (PHILOX): draw number `ctr` of stream `depth` of (pixel, launchSeed); no carried
// state besides the counter.
struct Rng {
  int mode = 0;
  int seed = 0;
  uint32_t pixel = 0, launchSeed = 0, depth = 0, ctr = 0;
};

inline Rng rngForPixel(int mode, uint32_t pixel, uint32_t launchSeed) {
  Rng r; r.mode = mode; r.pixel = pixel; r.launchSeed = launchSeed; r.depth = 1; r.ctr = 0;
  r.seed = (int)tea<16>(pixel, launchSeed);  // Camera.cu:24
  return r;
}

// 24-bit uniform integer; REF: utils_device.h:24-29.
inline unsigned int lcg(Rng& r) {
  if (r.mode == 0) {
    const unsigned int LCG_A = 1664525u, LCG_C = 1013904223u;
    r.seed = (int)(LCG_A * (unsigned int)r.seed + LCG_C);
    return (unsigned int)r.seed & 0x00FFFFFF;
  }
  uint32_t o[4];
  philox4x32_10(r.ctr >> 2, r.depth, 0u, 0u, r.pixel, r.launchSeed, o);
  uint32_t w = o[r.ctr & 3u];
  r.ctr++;
  return w >> 8;
}

inline float rnd(Rng& r) { return (float)lcg(r) / (float)0x01000000; }  // utils_device.h:32-34

// Child payload seed: folkPayload (utils_device.h:192-198) and the inline copies in
// Material.cu:60-63,97-101.  childDepth = parent depth + 1.
inline Rng forkRng(const Rng& parent, int childDepth) {
  Rng c = parent;
  c.depth = (uint32_t)childDepth;
  c.ctr = 0;
  c.seed = (int)tea<16>((unsigned int)parent.seed, (unsigned int)childDepth);
  return c;
}

inline float3 randInUnitSphere(Rng& r) {  // utils_device.h:36-43
  float3 p;
  do {
    float a = rnd(r), b = rnd(r), c = rnd(r);
    p = make_float3(a, b, c) * 2.0f - make_float3(1.f, 1.f, 1.f);
  } while (length(p) >= 1.0f);
  return p;
}

inline float3 randInUnitDisk(Rng& r) {  // utils_device.h:45-52
  float3 p;
  do {
    float a = rnd(r), b = rnd(r);
    p = make_float3(a, b, 0.f) * 2.0f - make_float3(1.f, 1.f, 0.f);
  } while (length(p) >= 1.f);
  return p;
}

// ---------------------------------------------------------------- dielectric Fresnel
inline float fresnel(float cosThetaI, float cosThetaT, float refIdx) {  // utils_device.h:63-67
  float rs = (cosThetaI - cosThetaT * refIdx) / (cosThetaI + refIdx * cosThetaT);
  float rp = (cosThetaI * refIdx - cosThetaT) / (cosThetaI * refIdx + cosThetaT);
  return 0.5f * (rs * rs + rp * rp);
}

// ---------------------------------------------------------------- hit-point refinement
inline float intersectPlane(const float3& origin, const float3& direction, const float3& normal,
                            const float3& point) {  // utils_device.h:72-79
  return -(dot(normal, origin - point)) / dot(normal, direction);
}

inline float offsetCoord(float h, float n) {  // one lane of utils_device.h:82-104
  const float epsilon = 1.0e-4f;
  const float off = 4096.0f * 2.0f;
  if ((float_as_int(h) & 0x7fffffff) < float_as_int(epsilon)) return h + epsilon * n;
  return int_as_float(float_as_int(h) + int(copysignf(off, h) * n));
}
inline float3 offsetPoint(const float3& hit, const float3& n) {
  return make_float3(offsetCoord(hit.x, n.x), offsetCoord(hit.y, n.y), offsetCoord(hit.z, n.z));
}

inline void refineHitpoint(const float3& original, const float3& direction, const float3& normal,
                           const float3& p, float3& back, float3& front) {  // utils_device.h:108-128
  float refined_t = intersectPlane(original, direction, normal, p);
  float3 refined = original + refined_t * direction;
  if (dot(direction, normal) > 0.0f) {
    back = offsetPoint(refined, normal);
    front = offsetPoint(refined, -normal);
  } else {
    back = offsetPoint(refined, -normal);
    front = offsetPoint(refined, normal);
  }
}

// ---------------------------------------------------------------- microfacet helpers
inline float square(float x) { return x * x; }

inline float GTR1(float NDotH, float a) {  // utils_device.h:130-137
  if (a >= 1.f) return (1.f / kPiF);
  float a2 = a * a;
  float t = 1.f + (a2 - 1.f) * NDotH * NDotH;
  return (a2 - 1.0f) / (kPiF * logf(a2) * t);
}
inline float GTR2(float NDotH, float a) {  // :139-143
  float a2 = a * a;
  float t = 1.f + (a2 - 1.f) * NDotH * NDotH;
  return a2 / (kPiF * t * t);
}
inline float GTR2Aniso(float NdotH, float HdotX, float HdotY, float ax, float ay) {  // :149-151
  return 1 / (kPiF * ax * ay * square(square(HdotX / ax) + square(HdotY / ay) + NdotH * NdotH));
}
inline float schlickFresnel(float u) {  // :153-157
  float m = clampf(1.f - u, 0.f, 1.f);
  float m2 = m * m;
  return m2 * m2 * m;
}
inline float smithGGgx(float NdotV, float alphaG) {  // :159-163
  float a = alphaG * alphaG;
  float b = NdotV * NdotV;
  return 1.f / (NdotV + sqrtf(a + b - a * b));
}
inline float smithGGgxAniso(float NdotV, float VdotX, float VdotY, float ax, float ay) {  // :165-167
  // The reference spells 1.0 / (... sqrt(...)): a double divide of float operands, which
  // rounds to the same float as the float divide.
  return (float)(1.0 / (double)(NdotV + sqrtf(square(VdotX * ax) + square(VdotY * ay) + square(NdotV))));
}
inline float3 srgb2lin(const float3& v) {  // :173-175
  return make_float3(powf(v.x, 2.2f), powf(v.y, 2.2f), powf(v.z, 2.2f));
}
inline float powerHeuristic(float a, float b) {  // :182-185
  float t = a * a;
  return t / (b * b + t);
}

// ---------------------------------------------------------------- Disney BRDF
// disney.h:9-30.  Draw order pinned: r0 (lobe), then (u1,u2) or (phi-draw, xi).
inline void disneySample(Rng& rng, const DisneyParams& mp, const float3& N, float3& L, const float3& V,
                         float3& H) {
  float diffuseRatio = 0.5f * (1.0f - mp.metallic);
  Onb onb(N);
  float r0 = rnd(rng);
  if (r0 < diffuseRatio) {
    float u1 = rnd(rng);
    float u2 = rnd(rng);
    cosine_sample_hemisphere(u1, u2, L);
    onb.inverse_transform(L);
    L = normalize(L);
    H = normalize(L + V);
  } else {
    float a = fmaxf(0.001f, mp.roughness);
    float phi = rnd(rng) * 2.0f * kPiF;
    float xi = rnd(rng);
    float cosTheta = sqrtf((1.f - xi) / (1.0f + (a * a - 1.f) * xi));
    float sinTheta = sqrtf(1.0f - (cosTheta * cosTheta));
    float sinPhi = sinf(phi);
    float cosPhi = cosf(phi);
    H = make_float3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
    onb.inverse_transform(H);
    L = normalize(2.0f * dot(V, H) * H - V);
    H = normalize(H);
  }
}

inline float disneyPdf(const DisneyParams& mp, const float3& N, const float3& L, const float3& V,
                       const float3& H) {  // disney.h:32-46
  (void)V;
  float diffuseRatio = 0.5f * (1.0f - mp.metallic);
  float specularAlpha = fmaxf(0.001f, mp.roughness);
  float clearcoatAlpha = lerp(0.1f, 0.001f, mp.clearcoatGloss);
  float specularRatio = 1.f - diffuseRatio;
  float cosTheta = fabsf(dot(N, H));
  float pdfGTR1 = GTR1(cosTheta, clearcoatAlpha) * cosTheta;
  float pdfGTR2 = GTR2(cosTheta, specularAlpha) * cosTheta;
  float ratio = 1.0f / (1.0f + mp.clearcoat);
  float pdfH = lerp(pdfGTR1, pdfGTR2, ratio);
  float pdfL = (float)((double)pdfH / (4.0 * (double)fabsf(dot(L, H))));  // "4.0 *" is double in the source
  float pdfDiff = fabsf(dot(N, L)) / kPiF;
  return diffuseRatio * pdfDiff + specularRatio * pdfL;
}

inline float3 disneyEval(const DisneyParams& mp, const float3& baseColor, const float3& N, const float3& L,
                         const float3& V, const float3& H) {  // disney.h:48-91
  Onb onb(N);
  float NdotL = dot(N, L);
  float NdotV = dot(N, V);
  float NdotH = dot(N, H);
  float LdotH = dot(L, H);
  float3 Cdlin = srgb2lin(baseColor);
  float Cdlum = dot(Cdlin, make_float3(0.3f, 0.6f, 0.1f));
  float3 Ctint = Cdlum > 0.f ? Cdlin / Cdlum : make_float3(1.f);
  float3 Cspec0 = lerp(mp.specular * 0.08f * lerp(make_float3(1.f), Ctint, mp.specularTint), Cdlin, mp.metallic);
  float3 Csheen = lerp(make_float3(1.f), Ctint, mp.sheenTint);

  float FL = schlickFresnel(NdotL);
  float FV = schlickFresnel(NdotV);
  float Fd90 = 0.5f + 2.f * LdotH * LdotH * mp.roughness;
  float Fd = lerp(1.f, Fd90, FL) * lerp(1.f, Fd90, FV);

  float Fss90 = LdotH * LdotH * mp.roughness;
  float Fss = lerp(1.0f, Fss90, FL) * lerp(1.0f, Fss90, FV);
  float ss = 1.25f * (Fss * (1.f / (NdotL + NdotV) - 0.5f) + 0.5f);

  float aspect = sqrtf(1 - mp.anisotropic * 0.9f);
  float ax = fmaxf(.001f, square(mp.roughness) / aspect);
  float ay = fmaxf(.001f, square(mp.roughness) * aspect);
  float3 X = normalize(onb.m_tangent);
  float3 Y = normalize(cross(N, X));
  float Ds = GTR2Aniso(NdotH, dot(H, X), dot(H, Y), ax, ay);
  float FH = schlickFresnel(LdotH);
  float3 Fs = lerp(Cspec0, make_float3(1.f), FH);
  float Gs = smithGGgxAniso(NdotL, dot(L, X), dot(L, Y), ax, ay) * smithGGgxAniso(NdotV, dot(V, X), dot(V, Y), ax, ay);
  float3 Fsheen = FH * mp.sheen * Csheen;
  float Dr = GTR1(NdotH, lerp(0.1f, 0.001f, mp.clearcoatGloss));
  float Fr = lerp(0.04f, 1.f, FH);
  float Gr = smithGGgx(NdotL, 0.25f) * smithGGgx(NdotV, 0.25f);
  float3 brdf = ((1.0f / kPiF) * lerp(Fd, ss, mp.subsurface) * Cdlin + Fsheen) * (1.0f - mp.metallic) +
                Gs * Fs * Ds + make_float3(0.25f * mp.clearcoat * Gr * Fr * Dr);
  return brdf;
}

}  // namespace orc
