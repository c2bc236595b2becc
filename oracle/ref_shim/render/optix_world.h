// TEST INFRASTRUCTURE.  Stand-in for OptiX 5.1's <optix_world.h> that lets the reference's DEVICE
// sources — utils_device.h, disney.h, Geometry.cu, Material.cu, Camera.cu, miss.cu, Exception.cu —
// compile unchanged with g++, from where they lie under /root/reference, into
// oracle/_ref/libref_render.so (Makefile target `ref`).  Nothing of the reference is copied.
//
// What is here:
//  * CUDA vector types and the optixu math helpers those files call.  optixu_math_namespace.h /
//    optixu_aabb_namespace.h (OptiX SDK 5.1.1) are NOT vendored in /root/reference; they are
//    restated here from the published header, independently of oracle/vecmath.h — at this
//    boundary the two restatements check each other, nothing more ("parity unpinned" for the
//    SDK helpers themselves).
//  * rtDeclareVariable / rtBuffer / RT_PROGRAM as thread-local globals, and rtTrace,
//    rtPotentialIntersection, rtReportIntersection, rtTerminateRay, rtTex2D as hooks that the
//    harness (ref_render.cpp) implements with a brute-force loop over all primitives.
//  * make_float3 / make_float2 / cosine_sample_hemisphere are function-like macros that expand
//    to BRACED initialisation: C++ evaluates a braced list left to right, function arguments in
//    unspecified order (g++: right to left, nvcc: left to right).  The reference draws random
//    numbers inside such argument lists (utils_device.h:40,49; disney.h:13; Camera.cu:29), so
//    this pins the order nvcc gives.  One expression cannot be pinned this way:
//    `light.u * rand(s) + light.v * rand(s)` (Material.cu:180) — g++ evaluates the right operand
//    first; tests account for it with the oracle's orc_set_quad_light_draw_order knob.
#pragma once
#include <math.h>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <type_traits>

// The float overloads must be the ones overload resolution finds (CUDA: pow(float,float) -> powf,
// abs(float) -> fabsf); with only the C double versions the results would differ in the last bit.
static_assert(std::is_same<decltype(pow(1.f, 2.2f)), float>::value, "pow(float,float) must be the float overload");
static_assert(std::is_same<decltype(abs(1.f)), float>::value, "abs(float) must be the float overload");
static_assert(std::is_same<decltype(sqrt(1.f)), float>::value, "sqrt(float) must be the float overload");
static_assert(std::is_same<decltype(copysign(1.f, 1.f)), float>::value, "copysign(float,float) must be the float overload");

#define __device__
#define __host__
#define __inline__ inline
#define RT_PROGRAM
#define RT_TEXTURE_ID_NULL 0
#define RT_DEFAULT_MAX 1.e27f
#ifndef M_PIf
#define M_PIf 3.14159265358979323846f
#endif

typedef unsigned int uint;
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int3 { int x, y, z; };
struct uint2 { unsigned int x, y; };
struct uchar4 { unsigned char x, y, z, w; };
typedef int rtObject;

inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float __saturatef(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline float min(float a, float b) { return fminf(a, b); }

namespace optix {
using ::float2; using ::float3; using ::float4; using ::int3; using ::uint2; using ::uchar4;

// Braced construction helpers behind the make_* macros (left-to-right evaluation).
struct mk2 : float2 {
  mk2(float a, float b) : float2{a, b} {}
  mk2(const uint2& u) : float2{(float)u.x, (float)u.y} {}
};
struct mk3 : float3 {
  mk3(float a, float b, float c) : float3{a, b, c} {}
  mk3(float s) : float3{s, s, s} {}
  mk3(const float4& v) : float3{v.x, v.y, v.z} {}
  mk3(const float2& v) : float3{v.x, v.y, 0.0f} {}
};
inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }

// ---- float2
inline float2 operator+(const float2& a, const float2& b) { return float2{a.x + b.x, a.y + b.y}; }
inline float2 operator-(const float2& a, float b) { return float2{a.x - b, a.y - b}; }
inline float2 operator*(const float2& a, float s) { return float2{a.x * s, a.y * s}; }
inline float2 operator/(const float2& a, const float2& b) { return float2{a.x / b.x, a.y / b.y}; }

// ---- float3
inline float3 operator+(const float3& a, const float3& b) { return float3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator+(const float3& a, float b) { return float3{a.x + b, a.y + b, a.z + b}; }
inline float3 operator-(const float3& a, const float3& b) { return float3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator-(const float3& a, float b) { return float3{a.x - b, a.y - b, a.z - b}; }
inline float3 operator-(const float3& a) { return float3{-a.x, -a.y, -a.z}; }
inline float3 operator*(const float3& a, const float3& b) { return float3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float3 operator*(const float3& a, float s) { return float3{a.x * s, a.y * s, a.z * s}; }
inline float3 operator*(float s, const float3& a) { return float3{s * a.x, s * a.y, s * a.z}; }
inline float3 operator/(const float3& a, const float3& b) { return float3{a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float3 operator/(const float3& a, float s) { float inv = 1.0f / s; return a * inv; }  // SDK: reciprocal, then multiply
inline void operator+=(float3& a, const float3& b) { a.x += b.x; a.y += b.y; a.z += b.z; }
inline void operator*=(float3& a, const float3& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }

inline float3 fminf(const float3& a, const float3& b) { return float3{::fminf(a.x, b.x), ::fminf(a.y, b.y), ::fminf(a.z, b.z)}; }
inline float3 fmaxf(const float3& a, const float3& b) { return float3{::fmaxf(a.x, b.x), ::fmaxf(a.y, b.y), ::fmaxf(a.z, b.z)}; }
using ::fminf; using ::fmaxf;

inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(const float3& a, const float3& b) {
  return float3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(const float3& v) { return sqrtf(dot(v, v)); }
inline float3 normalize(const float3& v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }

inline float lerp(float a, float b, float t) { return a + t * (b - a); }
inline float3 lerp(const float3& a, const float3& b, float t) { return a + t * (b - a); }
inline float clamp(float f, float a, float b) { return ::fmaxf(a, ::fminf(f, b)); }
inline float3 clamp(const float3& v, const float3& a, const float3& b) {
  return float3{clamp(v.x, a.x, b.x), clamp(v.y, a.y, b.y), clamp(v.z, a.z, b.z)};
}

inline float3 reflect(const float3& i, const float3& n) { return i - 2.0f * n * dot(n, i); }
inline float3 faceforward(const float3& n, const float3& i, const float3& nref) { return n * copysignf(1.0f, dot(i, nref)); }
inline bool refract(float3& r, const float3& i, const float3& n, float ior) {
  float3 nn = n;
  float negNdotV = dot(i, nn);
  float eta;
  if (negNdotV > 0.0f) { eta = ior; nn = -n; negNdotV = -negNdotV; }
  else { eta = 1.f / ior; }
  const float k = 1.f - eta * eta * (1.f - negNdotV * negNdotV);
  if (k < 0.0f) { r = float3{0.f, 0.f, 0.f}; return false; }
  r = normalize(eta * i - (eta * negNdotV + sqrtf(k)) * nn);
  return true;
}

struct Ray {
  Ray() {}
  Ray(const float3& o, const float3& d, unsigned int type, float tmin_, float tmax_ = RT_DEFAULT_MAX)
      : origin(o), direction(d), ray_type(type), tmin(tmin_), tmax(tmax_) {}
  float3 origin, direction;
  unsigned int ray_type;
  float tmin, tmax;
};

// optixu_math_namespace.h intersect_triangle_branchless: n = un-normalised CCW normal,
// beta <-> p1, gamma <-> p2, no culling.
inline bool intersect_triangle(const Ray& ray, const float3& p0, const float3& p1, const float3& p2, float3& n,
                               float& t, float& beta, float& gamma) {
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
  Onb(const float3& normal) {
    m_normal = normal;
    if (fabsf(m_normal.x) > fabsf(m_normal.z)) { m_binormal.x = -m_normal.y; m_binormal.y = m_normal.x; m_binormal.z = 0; }
    else { m_binormal.x = 0; m_binormal.y = -m_normal.z; m_binormal.z = m_normal.y; }
    m_binormal = normalize(m_binormal);
    m_tangent = cross(m_binormal, m_normal);
  }
  void inverse_transform(float3& p) const { p = p.x * m_tangent + p.y * m_binormal + p.z * m_normal; }
  float3 m_tangent, m_binormal, m_normal;
};

// Behind the cosine_sample_hemisphere macro: constructed from a braced list.
struct cosine_sample_hemisphere_call {
  cosine_sample_hemisphere_call(float u1, float u2, float3& p) {
    const float r = sqrtf(u1);
    const float phi = 2.0f * M_PIf * u2;
    p.x = r * cosf(phi);
    p.y = r * sinf(phi);
    p.z = sqrtf(::fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
  }
};

struct Aabb {
  float3 m_min, m_max;
  void set(const float3& mn, const float3& mx) { m_min = mn; m_max = mx; }
  void invalidate() { m_min = float3{1e37f, 1e37f, 1e37f}; m_max = float3{-1e37f, -1e37f, -1e37f}; }
};

template <class T> T rtTex2D(int id, float x, float y);
}  // namespace optix

#define make_float2(...) (::optix::mk2{__VA_ARGS__})
#define make_float3(...) (::optix::mk3{__VA_ARGS__})
#define cosine_sample_hemisphere(...) ((void)::optix::cosine_sample_hemisphere_call{__VA_ARGS__})

// ---- OptiX device-side objects as thread-local globals + hooks into the harness
namespace refshim {
template <class T, int D = 1>
struct Buffer {
  T* data = nullptr;
  size_t w = 0, h = 1;
  size_t size() const { return w; }
  T& operator[](size_t i) { return data[i]; }
  T& operator[](const uint2& i) { return data[(size_t)i.y * w + i.x]; }
};
void trace(const optix::Ray& ray, void* payload);
bool potentialIntersection(float t);
bool reportIntersection(unsigned int material);
void terminateRay();
float4 tex2D(int id, float x, float y);
}  // namespace refshim

#define rtDeclareVariable(type, name, ...) static thread_local type name
#define rtBuffer static thread_local ::refshim::Buffer

template <class P> inline void rtTrace(rtObject, const optix::Ray& ray, P& payload) { refshim::trace(ray, &payload); }
inline bool rtPotentialIntersection(float t) { return refshim::potentialIntersection(t); }
inline bool rtReportIntersection(unsigned int material) { return refshim::reportIntersection(material); }
inline void rtTerminateRay() { refshim::terminateRay(); }
namespace optix {
template <> inline float4 rtTex2D<float4>(int id, float x, float y) { return refshim::tex2D(id, x, y); }
}
