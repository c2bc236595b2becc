// TEST INFRASTRUCTURE.  Minimal stand-in for OptiX's <optix_world.h> so that the reference's
// scene.cpp / scene.h / Structures.h / tiny_obj_loader.h compile unchanged, from where they lie
// under /root/reference, into oracle/_ref/libref_loader.so (see Makefile target `ref`).  Only
// what those files touch: optix::float3/float4 with a few operators, length/cross/normalize
// (formulas of the SDK header), M_PIf, RT_TEXTURE_ID_NULL.
#pragma once
#include <cmath>
namespace optix {
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
inline float3 make_float3(float x, float y, float z) { float3 r = {x, y, z}; return r; }
inline float3 make_float3(float s) { return make_float3(s, s, s); }
inline float3 operator-(const float3& a, const float3& b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator+(const float3& a, const float3& b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator*(const float3& a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(const float3& a, const float3& b) {
  return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float length(const float3& v) { return sqrtf(dot(v, v)); }
inline float3 normalize(const float3& v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }
}  // namespace optix
#ifndef M_PIf
#define M_PIf 3.14159265358979323846f
#endif
#define RT_TEXTURE_ID_NULL 0
