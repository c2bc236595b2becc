// traverse.cuh — software ray traversal (what OptiX's rtTrace did inside the closed runtime;
// call sites Camera.cu:37, Material.cu:41,64,108,165,192,213) and the leaf primitive tests
// (Geometry.cu:18-55 sphere, :70-91 quad, :121-134 mesh via the SDK's intersect_triangle).
//
// Closest hit obeys the (t, primitive id) lexicographic rule; primitive tests use the exact
// operation sequence of the oracle (no FMA), so on identical rays the winning id, t, beta and
// gamma are bit-identical.  Box slabs are conservative ((lo - o) * 1/d, far side widened by
// 1e-5 relative) and may use any rounding.
#pragma once
#include "gpu_types.h"
#include "vec.cuh"

struct Hit { float t; int prim; float beta, gamma; };

// SDK intersect_triangle (branch-free variant) on a packed record.
MOX_D bool triTest(const float3& o, const float3& d, float tmin, const float3& p0, const float3& e0, const float3& e1,
                   float& t, float& beta, float& gamma) {
  const float3 n = cross(e1, e0);
  const float3 e2 = (1.0f / dot(n, d)) * (p0 - o);
  const float3 i = cross(d, e2);
  beta = dot(i, e1);
  gamma = dot(i, e0);
  t = dot(n, e2);
  return (t > tmin) & (beta >= 0.0f) & (gamma >= 0.0f) & (beta + gamma <= 1);
}

// First root in (tmin, bound) — near root, else far root (Geometry.cu:18-55).  `incl`: also accept t == bound.
MOX_D bool sphereTest(const float4& cr, const float3& o, const float3& d, float tmin, float bound, bool incl, float& t) {
  float3 oc = o - mk3(cr);
  float b = dot(d, oc);
  float c = dot(oc, oc) - cr.w * cr.w;
  float disc = b * b - c;
  if (disc < 0) return false;
  float root = sqrtf(disc);
  t = -b - root;
  if (t > tmin && (t < bound || (incl && t == bound))) return true;
  t = -b + root;
  return t > tmin && (t < bound || (incl && t == bound));
}

MOX_D bool quadTest(const Analytic& q, const float3& o, const float3& d, float tmin, float& t, float& a1, float& a2) {
  float3 n = mk3(q.a);
  float dt = dot(d, n);
  t = (q.a.w - dot(n, o)) / dt;
  if (!(t > tmin)) return false;
  float3 p = o + d * t;
  float3 vi = p - mk3(q.d);
  a1 = dot(mk3(q.b), vi);
  if (!(a1 >= 0 && a1 <= 1)) return false;
  a2 = dot(mk3(q.c), vi);
  return a2 >= 0 && a2 <= 1;
}

struct RayPre { float3 o, d, idir; float tmin; };

MOX_D RayPre prepRay(const float3& o, const float3& d, float tmin) {
  RayPre r;
  r.o = o; r.d = d; r.tmin = tmin;
  const float tiny = 1e-30f;
  r.idir.x = 1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x));
  r.idir.y = 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y));
  r.idir.z = 1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));
  return r;
}

// Entry distance of the slab test, or +inf when the box is missed within [tmin, tcur].
MOX_D float boxEntry(const RayPre& r, float lox, float hix, float loy, float hiy, float loz, float hiz, float tcur) {
  float x0 = (lox - r.o.x) * r.idir.x, x1 = (hix - r.o.x) * r.idir.x;
  float y0 = (loy - r.o.y) * r.idir.y, y1 = (hiy - r.o.y) * r.idir.y;
  float z0 = (loz - r.o.z) * r.idir.z, z1 = (hiz - r.o.z) * r.idir.z;
  float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), r.tmin));
  float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tcur)) * 1.00001f;
  return tn <= tf ? tn : __int_as_float(0x7f800000);
}

#define MOX_STACK 64

// Closest hit over the binary BVH.  COUNT adds node-visit / primitive-test counters.
template <bool COUNT>
MOX_D Hit traceClosest(const SceneView& s, const float3& o, const float3& d, float tmin, float tmax, uint32_t* nodeVisits,
                       uint32_t* primTests) {
  Hit best; best.t = tmax; best.prim = -1; best.beta = 0.f; best.gamma = 0.f;
  RayPre r = prepRay(o, d, tmin);
  int stack[MOX_STACK];
  int sp = 0;
  int cur = 0;  // root
  uint32_t nv = 0, np = 0;
  while (true) {
    if (cur >= 0) {
      const BvhNode2* nd = s.nodes + cur;
      float4 a = __ldg(&nd->c0xy), b = __ldg(&nd->c1xy), z = __ldg(&nd->cz);
      int4 ref = __ldg(&nd->ref);
      if (COUNT) nv++;
      float t0 = ref.x == MOX_EMPTY_CHILD ? __int_as_float(0x7f800000) : boxEntry(r, a.x, a.y, a.z, a.w, z.x, z.y, best.t);
      float t1 = ref.y == MOX_EMPTY_CHILD ? __int_as_float(0x7f800000) : boxEntry(r, b.x, b.y, b.z, b.w, z.z, z.w, best.t);
      bool h0 = t0 < __int_as_float(0x7f800000), h1 = t1 < __int_as_float(0x7f800000);
      if (h0 && h1) {
        bool swap = t1 < t0;
        int nearC = swap ? ref.y : ref.x, farC = swap ? ref.x : ref.y;
        if (sp < MOX_STACK) stack[sp++] = farC;
        cur = nearC;
        continue;
      } else if (h0 || h1) {
        cur = h0 ? ref.x : ref.y;
        continue;
      }
    } else {
      uint32_t leaf = (uint32_t)~cur;
      uint32_t first = leaf >> 3, count = (leaf & 7u) + 1u;
      for (uint32_t k = 0; k < count; ++k) {
        const float4* rec = s.packed + (size_t)(first + k) * MOX_PACKED_F4;
        float4 r0 = __ldg(rec);
        if (COUNT) np++;
        uint32_t idbits = __float_as_uint(r0.w);
        uint32_t type = idbits >> 30;
        int id = (int)(idbits & 0x3fffffffu);
        if (type == PT_TRI) {
          float4 r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
          float t, be, ga;
          if (triTest(o, d, tmin, mk3(r0), mk3(r1), mk3(r2), t, be, ga) && (t < best.t || (t == best.t && id < best.prim))) {
            best.t = t; best.prim = id; best.beta = be; best.gamma = ga;
          }
        } else {
          const Analytic* an = s.analytic + __float_as_int(r0.x);
          if (type == PT_SPHERE) {
            float t;
            if (sphereTest(__ldg(&an->a), o, d, tmin, best.t, id < best.prim, t)) { best.t = t; best.prim = id; best.beta = 0.f; best.gamma = 0.f; }
          } else {
            Analytic q;
            q.a = __ldg(&an->a); q.b = __ldg(&an->b); q.c = __ldg(&an->c); q.d = __ldg(&an->d);
            float t, a1, a2;
            if (quadTest(q, o, d, tmin, t, a1, a2) && (t < best.t || (t == best.t && id < best.prim))) {
              best.t = t; best.prim = id; best.beta = a1; best.gamma = a2;
            }
          }
        }
      }
    }
    if (sp == 0) break;
    cur = stack[--sp];
  }
  if (COUNT) { *nodeVisits = nv; *primTests = np; }
  return best;
}

// Shadow-ray transmittance with the order-independent rule (reference
// any-hit Material.cu:225-232): only Disney primitives occlude; any NORMAL hit in
// (tmin, tmax) -> 0, otherwise the product of the GLASS colours.
MOX_D float3 traceShadow(const SceneView& s, const float3& o, const float3& d, float tmin, float tmax) {
  float3 atten = mk3(1.f);
  RayPre r = prepRay(o, d, tmin);
  int stack[MOX_STACK];
  int sp = 0;
  int cur = 0;
  while (true) {
    if (cur >= 0) {
      const BvhNode2* nd = s.nodes + cur;
      float4 a = __ldg(&nd->c0xy), b = __ldg(&nd->c1xy), z = __ldg(&nd->cz);
      int4 ref = __ldg(&nd->ref);
      bool h0 = ref.x != MOX_EMPTY_CHILD && boxEntry(r, a.x, a.y, a.z, a.w, z.x, z.y, tmax) < __int_as_float(0x7f800000);
      bool h1 = ref.y != MOX_EMPTY_CHILD && boxEntry(r, b.x, b.y, b.z, b.w, z.z, z.w, tmax) < __int_as_float(0x7f800000);
      if (h0 && h1) { if (sp < MOX_STACK) stack[sp++] = ref.y; cur = ref.x; continue; }
      else if (h0 || h1) { cur = h0 ? ref.x : ref.y; continue; }
    } else {
      uint32_t leaf = (uint32_t)~cur;
      uint32_t first = leaf >> 3, count = (leaf & 7u) + 1u;
      for (uint32_t k = 0; k < count; ++k) {
        const float4* rec = s.packed + (size_t)(first + k) * MOX_PACKED_F4;
        float4 r0 = __ldg(rec);
        uint32_t idbits = __float_as_uint(r0.w);
        uint32_t type = idbits >> 30;
        uint32_t id = idbits & 0x3fffffffu;
        const GpuMaterial* m = s.mats + (__ldg(&s.prims[id].typeMat) >> 2);
        if (__ldg(&m->kind) != MOX_MAT_DISNEY) continue;
        bool hit;
        if (type == PT_TRI) {
          float4 r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
          float t, be, ga;
          hit = triTest(o, d, tmin, mk3(r0), mk3(r1), mk3(r2), t, be, ga) && t < tmax;
        } else {
          const Analytic* an = s.analytic + __float_as_int(r0.x);
          if (type == PT_SPHERE) {
            float t;
            hit = sphereTest(__ldg(&an->a), o, d, tmin, tmax, false, t);
          } else {
            Analytic q;
            q.a = __ldg(&an->a); q.b = __ldg(&an->b); q.c = __ldg(&an->c); q.d = __ldg(&an->d);
            float t, a1, a2;
            hit = quadTest(q, o, d, tmin, t, a1, a2) && t < tmax;
          }
        }
        if (hit) {
          if (m->dis.brdfType == GLASS) atten *= f3(m->dis.color);
          else return mk3(0.f);
        }
      }
    }
    if (sp == 0) break;
    cur = stack[--sp];
  }
  return atten;
}
