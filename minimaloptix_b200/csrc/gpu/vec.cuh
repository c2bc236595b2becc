// vec.cuh — float3 arithmetic for the CUDA kernels.
//
// The whole library is compiled with -fmad=false: every a*b+c written with operators rounds
// twice, exactly like the CPU oracle (g++ -ffp-contract=off), so intersection results
// (t, beta, gamma, primitive id) are bit-identical on identical rays and shading follows the
// same rounding sequence.  Where a fused multiply-add is wanted for speed and does not affect
// parity (conservative ray/box slabs) it is spelled explicitly with fmaf().
// Operator semantics follow the OptiX SDK helpers the reference device code is written
// against (normalize = v * (1/sqrt(dot)); float3/float = multiply by reciprocal), restated in
// SURVEY.md §8(c).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mox_structs.h"

#define MOX_HD __host__ __device__ __forceinline__
#define MOX_D __device__ __forceinline__

MOX_HD float3 mk3(float x, float y, float z) { return make_float3(x, y, z); }
MOX_HD float3 mk3(float s) { return make_float3(s, s, s); }
MOX_HD float3 mk3(const float4& v) { return make_float3(v.x, v.y, v.z); }
MOX_HD float3 f3(const mox_float3& v) { return make_float3(v.x, v.y, v.z); }
MOX_HD float3 operator+(const float3& a, const float3& b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
MOX_HD float3 operator-(const float3& a, const float3& b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
MOX_HD float3 operator-(const float3& a) { return mk3(-a.x, -a.y, -a.z); }
MOX_HD float3 operator*(const float3& a, const float3& b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
MOX_HD float3 operator*(const float3& a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
MOX_HD float3 operator*(float s, const float3& a) { return mk3(s * a.x, s * a.y, s * a.z); }
MOX_HD float3 operator/(const float3& a, float s) { float inv = 1.0f / s; return a * inv; }
MOX_HD float3 operator+(const float3& a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
MOX_HD float3& operator+=(float3& a, const float3& b) { a = a + b; return a; }
MOX_HD float3& operator*=(float3& a, const float3& b) { a = a * b; return a; }
MOX_HD float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
MOX_HD float3 cross(const float3& a, const float3& b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
MOX_HD float length(const float3& v) { return sqrtf(dot(v, v)); }
MOX_HD float3 normalize(const float3& v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }
MOX_HD float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
MOX_HD float lerpf(float a, float b, float t) { return a + t * (b - a); }
MOX_HD float3 lerp3(const float3& a, const float3& b, float t) { return a + t * (b - a); }
MOX_HD float3 fmin3(const float3& a, const float3& b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
MOX_HD float3 fmax3(const float3& a, const float3& b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
MOX_HD float3 reflect3(const float3& i, const float3& n) { return i - 2.0f * n * dot(n, i); }
MOX_HD float3 faceforward3(const float3& n, const float3& i, const float3& nref) { return n * copysignf(1.0f, dot(i, nref)); }

// SDK refract(): false on total internal reflection.
MOX_HD bool refract3(float3& r, const float3& i, const float3& n, float ior) {
  float3 nn = n;
  float negNdotV = dot(i, nn);
  float eta;
  if (negNdotV > 0.0f) { eta = ior; nn = -n; negNdotV = -negNdotV; }
  else { eta = 1.f / ior; }
  const float k = 1.f - eta * eta * (1.f - negNdotV * negNdotV);
  if (k < 0.0f) { r = mk3(0.f); return false; }
  r = normalize(eta * i - (eta * negNdotV + sqrtf(k)) * nn);
  return true;
}

// Orthonormal basis around a unit normal (SDK Onb).
struct Onb3 {
  float3 tangent, binormal, normal;
  MOX_HD explicit Onb3(const float3& n) {
    normal = n;
    if (fabsf(n.x) > fabsf(n.z)) binormal = mk3(-n.y, n.x, 0.f);
    else binormal = mk3(0.f, -n.z, n.y);
    binormal = normalize(binormal);
    tangent = cross(binormal, normal);
  }
  MOX_HD float3 toWorld(const float3& p) const { return p.x * tangent + p.y * binormal + p.z * normal; }
};

#define MOX_PI_F 3.14159265358979323846f
#define MOX_RAY_TMAX 1.e27f
