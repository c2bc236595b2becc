// oracle/vecmath.h — TEST INFRASTRUCTURE (CPU oracle), never linked into the product.
//
// Scalar restatement of the OptiX SDK 5.1.1 math helpers the reference device code
// calls (optixu/optixu_math_namespace.h — NOT vendored in /root/reference; restated
// from the published header; "parity unpinned" at this boundary:
// the reference ships no tests or vectors for them).  Call sites in the reference:
// Geometry.cu:133 (intersect_triangle), Material.cu:56,90,103,125 (reflect, refract,
// faceforward), disney.h:11,13 (Onb, cosine_sample_hemisphere), disney.h:37,57 (lerp).
//
// Compile with -ffp-contract=off so that every a*b+c is two IEEE roundings, the same
// sequence the CUDA side spells with __fmul_rn/__fadd_rn in its intersection code.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };

inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
inline float3 make_float3(float s) { return {s, s, s}; }
inline float3 make_float3(const float4& v) { return {v.x, v.y, v.z}; }

inline float3 operator+(const float3& a, const float3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator-(const float3& a, const float3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator-(const float3& a) { return {-a.x, -a.y, -a.z}; }
inline float3 operator*(const float3& a, const float3& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float3 operator*(const float3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator*(float s, const float3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline float3 operator/(const float3& a, const float3& b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
// SDK: operator/(float3, float) multiplies by the reciprocal.
inline float3 operator/(const float3& a, float s) { float inv = 1.0f / s; return a * inv; }
inline float3 operator+(const float3& a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline float3 operator-(const float3& a, float s) { return {a.x - s, a.y - s, a.z - s}; }
inline float3& operator+=(float3& a, const float3& b) { a = a + b; return a; }
inline float3& operator*=(float3& a, const float3& b) { a = a * b; return a; }

inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(const float3& a, const float3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(const float3& v) { return sqrtf(dot(v, v)); }
inline float3 normalize(const float3& v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }

inline float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
inline float3 clamp(const float3& v, const float3& a, const float3& b) {
  return {clampf(v.x, a.x, b.x), clampf(v.y, a.y, b.y), clampf(v.z, a.z, b.z)};
}
inline float lerp(float a, float b, float t) { return a + t * (b - a); }
inline float3 lerp(const float3& a, const float3& b, float t) { return a + t * (b - a); }
inline float3 fminf3(const float3& a, const float3& b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
inline float3 fmaxf3(const float3& a, const float3& b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }

inline float3 reflect(const float3& i, const float3& n) { return i - 2.0f * n * dot(n, i); }
inline float3 faceforward(const float3& n, const float3& i, const float3& nref) {
  return n * copysignf(1.0f, dot(i, nref));
}
// Returns false on total internal reflection.
inline bool refract(float3& r, const float3& i, const float3& n, float ior) {
  float3 nn = n;
  float negNdotV = dot(i, nn);
  float eta;
  if (negNdotV > 0.0f) { eta = ior; nn = -n; negNdotV = -negNdotV; }
  else { eta = 1.f / ior; }
  const float k = 1.f - eta * eta * (1.f - negNdotV * negNdotV);
  if (k < 0.0f) { r = make_float3(0.f); return false; }
  r = normalize(eta * i - (eta * negNdotV + sqrtf(k)) * nn);
  return true;
}

struct Ray { float3 origin; float3 direction; float tmin; float tmax; };
constexpr float RT_DEFAULT_MAX = 1.e27f;

// The SDK's branch-free triangle test: n is the un-normalised CCW normal, beta <-> p1,
// gamma <-> p2, two-sided.
inline bool intersect_triangle(const Ray& ray, const float3& p0, const float3& p1, const float3& p2,
                               float3& n, float& t, float& beta, float& gamma) {
  const float3 e0 = p1 - p0;
  const float3 e1 = p0 - p2;
  n = cross(e1, e0);
  const float3 e2 = (1.0f / dot(n, ray.direction)) * (p0 - ray.origin);
  const float3 i = cross(ray.direction, e2);
  beta = dot(i, e1);
  gamma = dot(i, e0);
  t = dot(n, e2);
  return ((t < ray.tmax) & (t > ray.tmin) & (beta >= 0.0f) & (gamma >= 0.0f) & (beta + gamma <= 1));
}

struct Onb {
  explicit Onb(const float3& normal) {
    m_normal = normal;
    if (fabsf(m_normal.x) > fabsf(m_normal.z)) {
      m_binormal.x = -m_normal.y; m_binormal.y = m_normal.x; m_binormal.z = 0;
    } else {
      m_binormal.x = 0; m_binormal.y = -m_normal.z; m_binormal.z = m_normal.y;
    }
    m_binormal = normalize(m_binormal);
    m_tangent = cross(m_binormal, m_normal);
  }
  void inverse_transform(float3& p) const { p = p.x * m_tangent + p.y * m_binormal + p.z * m_normal; }
  float3 m_tangent, m_binormal, m_normal;
};

constexpr float kPiF = 3.14159265358979323846f;

inline void cosine_sample_hemisphere(float u1, float u2, float3& p) {
  const float r = sqrtf(u1);
  const float phi = 2.0f * kPiF * u2;
  p.x = r * cosf(phi);
  p.y = r * sinf(phi);
  p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
}

inline int float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline float int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }

}  // namespace orc
