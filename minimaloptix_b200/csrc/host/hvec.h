// hvec.h — small float3 helpers for the host side (operate on the ABI's mox_float3).
// Formulas follow the OptiX SDK helpers the reference host code uses (normalize multiplies
// by 1/sqrt(dot); float3/float multiplies by the reciprocal) so CamParams / QuadParams /
// LightParams come out as the reference computes them (utils_host.cpp:67-99, scene.cpp:78-88).
#pragma once
#include <cmath>
#include "mox_structs.h"

namespace moxh {
typedef mox_float3 float3;
inline float3 mk3(float x, float y, float z) { float3 r; r.x = x; r.y = y; r.z = z; return r; }
inline float3 mk3(float s) { return mk3(s, s, s); }
inline float3 operator+(float3 a, float3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(float3 a, float3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator-(float3 a) { return mk3(-a.x, -a.y, -a.z); }
inline float3 operator*(float3 a, float3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator*(float3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float s, float3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
inline float3 operator/(float3 a, float s) { float inv = 1.0f / s; return a * inv; }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float length(float3 v) { return sqrtf(dot(v, v)); }
inline float3 normalize(float3 v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }
inline float3 vmin(float3 a, float3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
inline float3 vmax(float3 a, float3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }

struct Aabb {
  float3 lo, hi;
  Aabb() { invalidate(); }
  void invalidate() { lo = mk3(1e37f); hi = mk3(-1e37f); }
  void include(float3 p) { lo = vmin(lo, p); hi = vmax(hi, p); }
  float3 center() const { return (lo + hi) * 0.5f; }
  float3 extent() const { return hi - lo; }
  bool valid() const { return lo.x <= hi.x && lo.y <= hi.y && lo.z <= hi.z; }
};
}  // namespace moxh
