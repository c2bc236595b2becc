// TEST INFRASTRUCTURE.  Harness around the REFERENCE's own device programs.  Camera.cu,
// Geometry.cu, Material.cu, miss.cu, Exception.cu, disney.h and utils_device.h are #included
// from /root/reference (never copied) and compiled with g++ against the shim
// oracle/ref_shim/render/optix_world.h into oracle/_ref/libref_render.so.  The closed part of
// OptiX (rtTrace traversal, program dispatch) is replaced by the simplest thing with the
// same semantics: every primitive is offered to its intersection program in primitive-id
// order ("NoAccel", MinimalOptiX.cpp:248), rtPotentialIntersection is the open interval
// (tmin, current tmax), rtReportIntersection runs the material's any-hit program for the ray
// type (only `disneyAnyHit` on shadow rays, MinimalOptiX.cpp:482-485) and accepts the hit,
// rtTerminateRay ends the loop, then closest-hit or miss runs for radiance rays.
//
// It exports (a) the render C ABI of include/mox.h under the prefix ref_ so the same scene
// upload code drives it, the oracle and the GPU library, and (b) the device helpers one by
// one.  scripts/make_render_golden.py uses it in this container to write
// tests/golden/render_ref.json; tests/test_ref_render.py compares oracle/ against those
// vectors bit for bit and, when this library is present, against the library directly.
//
// Semantics kept from OptiX / the reference, which the product deliberately changes
// (SURVEY.md App. C): primitives whose bounding-box PROGRAM yields an invalid box are never
// intersected (quadBBox / meshBBox); spheres are always offered (the reference's inverted
// sphere box, Geometry.cu:57-63, is harmless under NoAccel — Q7); a shadow ray's accepted
// GLASS hit shrinks tmax, so shadow results can depend on primitive order (Q8) — goldens use
// scenes where they do not.
#include <optix_world.h>
#include "structures.h"
#include "Structures.h"
#include "utils_device.h"
#include "disney.h"

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

// Status codes, material kinds and mox_stats come from the product's ABI header; its parameter
// structs are skipped — here Payload, CamParams, ... ARE the reference's Structures.h.
#define MOX_STRUCTS_H
#include "../../include/mox.h"

static_assert(sizeof(::Payload) == 32 && sizeof(::CamParams) == 76 && sizeof(::SphereParams) == 28, "layout");

namespace ref_geo {
#include "Geometry.cu"
}
namespace ref_mat {
#include "Material.cu"
}
namespace ref_cam {
#include "Camera.cu"
}
namespace ref_miss {
#include "miss.cu"
}
namespace ref_exc {
#include "Exception.cu"
}

namespace {

using optix::Ray;
enum PrimType { PT_SPHERE = 0, PT_QUAD = 1, PT_TRI = 2 };

struct Mesh {
  std::vector<float3> v, n;
  std::vector<float2> uv;
  std::vector<int3> vi, ni, ti;
};
struct Material {
  int kind;
  ::LambertianParams lam; ::MetalParams met; ::GlassParams gls; ::DisneyParams dis; ::LightParams lgt;
};
struct Prim { int type; int geom; int face; int mat; bool valid; };
struct Texture { int w, h; std::vector<float> texels; };
struct Attr { float3 geoNormal, shadingNormal, front, back, texcoord; };

}  // namespace

struct ref_ctx {
  std::string err;
  uint32_t W = 0, H = 0, maxDepth = 256;
  float eps = 0.001f, minIntensity = 0.001f;
  float3 absorb{0, 0, 0}, bad{1, 1, 1}, bg{0, 0, 0};
  ::CamParams cam{};
  int nThreads = 0;
  bool built = false;
  std::vector<Prim> prims;
  std::vector<::SphereParams> spheres;
  std::vector<::QuadParams> quads;
  std::vector<Mesh> meshes;
  std::vector<Material> mats;
  std::vector<::LightParams> lights;
  std::vector<Texture> textures;
  std::vector<float3> accu;
  uint64_t launches = 0;
  std::atomic<uint64_t> primary{0}, bounce{0}, shadow{0};
};

namespace {

// One rtTrace in flight (they nest: closest-hit programs call rtTrace).
struct TraceState {
  const ref_ctx* c;
  Ray ray;
  bool shadowRay;
  void* payload;
  float tmaxCur;
  int curPrim;
  float pendingT;
  bool haveHit = false, terminated = false;
  int hitPrim = -1;
  float hitT = 0;
  Attr hitAttr;
};
thread_local TraceState* g_trace = nullptr;
thread_local const ref_ctx* g_ctx = nullptr;

// All rtVariables of Material.cu that a nested trace overwrites.
struct MatGlobals {
  ::Payload payload; Ray ray; float t; Attr a;
  ::LambertianParams lam; ::MetalParams met; ::GlassParams gls; ::DisneyParams dis; ::LightParams lgt;
};
void saveMat(MatGlobals& g) {
  g.payload = ref_mat::payload; g.ray = ref_mat::ray; g.t = ref_mat::t;
  g.a = {ref_mat::geoNormal, ref_mat::shadingNormal, ref_mat::frontHitPoint, ref_mat::backHitPoint, ref_mat::texcoord};
  g.lam = ref_mat::lambParams; g.met = ref_mat::metalParams; g.gls = ref_mat::glassParams;
  g.dis = ref_mat::disneyParams; g.lgt = ref_mat::lightParams;
}
void restoreMat(const MatGlobals& g) {
  ref_mat::payload = g.payload; ref_mat::ray = g.ray; ref_mat::t = g.t;
  ref_mat::geoNormal = g.a.geoNormal; ref_mat::shadingNormal = g.a.shadingNormal; ref_mat::frontHitPoint = g.a.front;
  ref_mat::backHitPoint = g.a.back; ref_mat::texcoord = g.a.texcoord;
  ref_mat::lambParams = g.lam; ref_mat::metalParams = g.met; ref_mat::glassParams = g.gls;
  ref_mat::disneyParams = g.dis; ref_mat::lightParams = g.lgt;
}

void bindContext(const ref_ctx* c) {
  g_ctx = c;
  ref_mat::rayMaxDepth = c->maxDepth; ref_mat::rayTypeRadiance = 0; ref_mat::rayTypeShadow = 1;
  ref_mat::rayEpsilonT = c->eps; ref_mat::rayMinIntensity = c->minIntensity; ref_mat::absorbColor = c->absorb;
  ref_mat::topGroup = 0;
  ref_mat::lights.data = const_cast<::LightParams*>(c->lights.data()); ref_mat::lights.w = c->lights.size();
  ref_cam::rayTypeRadiance = 0; ref_cam::rayEpsilonT = c->eps; ref_cam::camParams = c->cam; ref_cam::topGroup = 0;
  ref_cam::launchDim = uint2{c->W, c->H}; ref_cam::nSuperSampling = 0;
  ref_cam::accuBuffer.data = const_cast<float3*>(c->accu.data()); ref_cam::accuBuffer.w = c->W; ref_cam::accuBuffer.h = c->H;
  ref_miss::bgColor = c->bg;
  ref_exc::badColor = c->bad;
  ref_exc::accuBuffer.data = const_cast<float3*>(c->accu.data()); ref_exc::accuBuffer.w = c->W; ref_exc::accuBuffer.h = c->H;
}

// Offer primitive `id` to its intersection program (Geometry.cu).
void intersectPrim(const ref_ctx& c, int id, const Ray& ray) {
  const Prim& p = c.prims[id];
  ref_geo::ray = ray;
  if (p.type == PT_SPHERE) {
    ref_geo::sphereParams = c.spheres[p.geom];
    ref_geo::sphereIntersect(0);
  } else if (p.type == PT_QUAD) {
    ref_geo::quadParams = c.quads[p.geom];
    ref_geo::quadIntersect(0);
  } else {
    const Mesh& m = c.meshes[p.geom];
    ref_geo::vertexBuffer.data = const_cast<float3*>(m.v.data()); ref_geo::vertexBuffer.w = m.v.size();
    ref_geo::normalBuffer.data = const_cast<float3*>(m.n.data()); ref_geo::normalBuffer.w = m.n.size();
    ref_geo::texcoordBuffer.data = const_cast<float2*>(m.uv.data()); ref_geo::texcoordBuffer.w = m.uv.size();
    ref_geo::vertIdxBuffer.data = const_cast<int3*>(m.vi.data()); ref_geo::vertIdxBuffer.w = m.vi.size();
    ref_geo::normIdxBuffer.data = const_cast<int3*>(m.ni.data()); ref_geo::normIdxBuffer.w = m.ni.size();
    ref_geo::texIdxBuffer.data = const_cast<int3*>(m.ti.data()); ref_geo::texIdxBuffer.w = m.ti.size();
    ref_geo::meshIntersect(p.face);
  }
}

void setMaterialParams(const Material& m) {
  switch (m.kind) {
    case MOX_MAT_LAMBERTIAN: ref_mat::lambParams = m.lam; break;
    case MOX_MAT_METAL: ref_mat::metalParams = m.met; break;
    case MOX_MAT_GLASS: ref_mat::glassParams = m.gls; break;
    case MOX_MAT_DISNEY: ref_mat::disneyParams = m.dis; break;
    case MOX_MAT_LIGHT: ref_mat::lightParams = m.lgt; break;
  }
}

// Bounding-box programs (Geometry.cu:57-63,93-110,162-175).
void primBounds(const ref_ctx& c, const Prim& p, float out[6]) {
  if (p.type == PT_SPHERE) {
    ref_geo::sphereParams = c.spheres[p.geom];
    ref_geo::sphereBBox(0, out);
  } else if (p.type == PT_QUAD) {
    ref_geo::quadParams = c.quads[p.geom];
    ref_geo::quadBBox(0, out);
  } else {
    const Mesh& m = c.meshes[p.geom];
    ref_geo::vertexBuffer.data = const_cast<float3*>(m.v.data()); ref_geo::vertexBuffer.w = m.v.size();
    ref_geo::vertIdxBuffer.data = const_cast<int3*>(m.vi.data()); ref_geo::vertIdxBuffer.w = m.vi.size();
    ref_geo::meshBBox(p.face, out);
  }
}

}  // namespace

// ---------------------------------------------------------------- the hooks the shim declares
namespace refshim {

bool potentialIntersection(float t) {
  TraceState& s = *g_trace;
  if (t > s.ray.tmin && t < s.tmaxCur) { s.pendingT = t; return true; }
  return false;
}

bool reportIntersection(unsigned int) {
  TraceState& s = *g_trace;
  const ref_ctx& c = *s.c;
  const Material& m = c.mats[c.prims[s.curPrim].mat];
  if (s.shadowRay && m.kind == MOX_MAT_DISNEY) {
    // any-hit program of the Disney material on the shadow ray type (Material.cu:225-232)
    ref_mat::disneyParams = m.dis;
    ref_mat::payload = *(::Payload*)s.payload;
    ref_mat::disneyAnyHit();
    *(::Payload*)s.payload = ref_mat::payload;
  }
  s.haveHit = true;
  s.hitPrim = s.curPrim;
  s.hitT = s.pendingT;
  s.tmaxCur = s.pendingT;
  s.hitAttr = {ref_geo::geoNormal, ref_geo::shadingNormal, ref_geo::frontHitPoint, ref_geo::backHitPoint, ref_geo::texcoord};
  return true;
}

void terminateRay() { g_trace->terminated = true; }

// rtTex2D<float4>: bilinear, texel centres at +0.5, normalized coordinates, REPEAT
// (MinimalOptiX.cpp:449-474).  Hardware behaviour restated; not reference text.
float4 tex2D(int id, float u, float v) {
  const Texture& tx = g_ctx->textures[id - 1];
  float x = u * tx.w - 0.5f, y = v * tx.h - 0.5f;
  float fx = floorf(x), fy = floorf(y);
  float ax = x - fx, ay = y - fy;
  auto wrap = [](int i, int n) { int m = i % n; return m < 0 ? m + n : m; };
  int x0 = wrap((int)fx, tx.w), x1 = wrap((int)fx + 1, tx.w);
  int y0 = wrap((int)fy, tx.h), y1 = wrap((int)fy + 1, tx.h);
  auto px = [&](int xi, int yi, int k) { return tx.texels[4 * ((size_t)yi * tx.w + xi) + k]; };
  float o[4];
  for (int k = 0; k < 4; ++k)
    o[k] = (px(x0, y0, k) * (1 - ax) + px(x1, y0, k) * ax) * (1 - ay) + (px(x0, y1, k) * (1 - ax) + px(x1, y1, k) * ax) * ay;
  return float4{o[0], o[1], o[2], o[3]};
}

void trace(const Ray& ray, void* payload) {
  const ref_ctx& c = *g_ctx;
  TraceState st;
  st.c = &c; st.ray = ray; st.shadowRay = ray.ray_type == 1; st.payload = payload; st.tmaxCur = ray.tmax;
  TraceState* outer = g_trace;
  MatGlobals saved;
  saveMat(saved);
  g_trace = &st;
  ref_ctx& mc = const_cast<ref_ctx&>(c);
  if (st.shadowRay) mc.shadow++;
  else if (((::Payload*)payload)->depth == 1) mc.primary++;
  else mc.bounce++;

  for (int id = 0; id < (int)c.prims.size() && !st.terminated; ++id) {
    if (!c.prims[id].valid) continue;
    st.curPrim = id;
    intersectPrim(c, id, ray);
  }
  if (!st.shadowRay) {
    if (st.haveHit) {
      const Material& m = c.mats[c.prims[st.hitPrim].mat];
      setMaterialParams(m);
      ref_mat::ray = ray; ref_mat::t = st.hitT;
      ref_mat::geoNormal = st.hitAttr.geoNormal; ref_mat::shadingNormal = st.hitAttr.shadingNormal;
      ref_mat::frontHitPoint = st.hitAttr.front; ref_mat::backHitPoint = st.hitAttr.back; ref_mat::texcoord = st.hitAttr.texcoord;
      ref_mat::payload = *(::Payload*)payload;
      switch (m.kind) {
        case MOX_MAT_LAMBERTIAN: ref_mat::lambertian(); break;
        case MOX_MAT_METAL: ref_mat::metal(); break;
        case MOX_MAT_GLASS: ref_mat::glass(); break;
        case MOX_MAT_DISNEY: ref_mat::disney(); break;
        case MOX_MAT_LIGHT: ref_mat::light(); break;
      }
      *(::Payload*)payload = ref_mat::payload;
    } else {
      ref_miss::ray = ray; ref_miss::pld = *(::Payload*)payload;
      ref_miss::staticMiss();  // only ray type 0 has a miss program (MinimalOptiX.cpp:164)
      *(::Payload*)payload = ref_miss::pld;
    }
  }
  restoreMat(saved);
  g_trace = outer;
}

}  // namespace refshim

namespace {

int addMaterial(ref_ctx* c, int kind, const void* params) {
  Material m{};
  m.kind = kind;
  switch (kind) {
    case MOX_MAT_LAMBERTIAN: memcpy(&m.lam, params, sizeof m.lam); break;
    case MOX_MAT_METAL: memcpy(&m.met, params, sizeof m.met); break;
    case MOX_MAT_GLASS: memcpy(&m.gls, params, sizeof m.gls); break;
    case MOX_MAT_DISNEY: memcpy(&m.dis, params, sizeof m.dis); break;
    case MOX_MAT_LIGHT: memcpy(&m.lgt, params, sizeof m.lgt); break;
    default: return -1;
  }
  c->mats.push_back(m);
  return (int)c->mats.size() - 1;
}

// One launch: camera() for every pixel (Camera.cu:21-42), rows split over threads.
void launchOnce(ref_ctx* c, int32_t seed) {
  int nt = c->nThreads > 0 ? c->nThreads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  std::atomic<uint32_t> next{0};
  auto worker = [&]() {
    bindContext(c);
    ref_cam::randSeed = seed;
    for (;;) {
      uint32_t y = next.fetch_add(1);
      if (y >= c->H) break;
      for (uint32_t x = 0; x < c->W; ++x) {
        ref_cam::launchIdx = uint2{x, y};
        ref_cam::camera();
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nt; ++i) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  c->launches++;
}

float3 a3(const float* p) { return float3{p[0], p[1], p[2]}; }
void s3(float* p, const float3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

}  // namespace

// ====================================================================== C ABI (prefix ref_)
extern "C" {

int ref_abi_version(void) { return MOX_ABI_VERSION; }
int ref_create(ref_ctx** out, int) { if (!out) return MOX_ERR_INVALID; *out = new ref_ctx(); return MOX_OK; }
void ref_destroy(ref_ctx* c) { delete c; }
const char* ref_last_error(const ref_ctx* c) { return c ? c->err.c_str() : ""; }

int ref_set_globals(ref_ctx* c, uint32_t w, uint32_t h, uint32_t maxDepth, float eps, float minI, const float absorb[3],
                    const float bad[3], const float bg[3]) {
  if (!c || w == 0 || h == 0) return MOX_ERR_INVALID;
  if (w != c->W || h != c->H) { c->accu.assign((size_t)w * h, float3{0, 0, 0}); c->launches = 0; }
  c->W = w; c->H = h; c->maxDepth = maxDepth; c->eps = eps; c->minIntensity = minI;
  c->absorb = a3(absorb); c->bad = a3(bad); c->bg = a3(bg);
  return MOX_OK;
}
int ref_set_camera(ref_ctx* c, const void* p) { if (!c || !p) return MOX_ERR_INVALID; memcpy(&c->cam, p, sizeof c->cam); return MOX_OK; }
int ref_set_rng_mode(ref_ctx* c, int m) {
  if (!c) return MOX_ERR_INVALID;
  if (m != 0) { c->err = "the reference has only the tea/lcg generator"; return MOX_ERR_INVALID; }
  return MOX_OK;
}
int ref_set_partition(ref_ctx* c, uint32_t rank, uint32_t world, uint32_t) {
  if (!c) return MOX_ERR_INVALID;
  if (rank != 0 || world != 1) { c->err = "the reference is single-device"; return MOX_ERR_INVALID; }
  return MOX_OK;
}
int ref_set_threads(ref_ctx* c, int n) { if (!c) return MOX_ERR_INVALID; c->nThreads = n; return MOX_OK; }

int ref_add_texture_rgba32f(ref_ctx* c, const float* texels, int w, int h, int* out_id) {
  if (!c || !texels || w <= 0 || h <= 0) return MOX_ERR_INVALID;
  Texture t; t.w = w; t.h = h; t.texels.assign(texels, texels + (size_t)w * h * 4);
  c->textures.push_back(std::move(t));
  if (out_id) *out_id = (int)c->textures.size();
  return MOX_OK;
}
int ref_add_sphere(ref_ctx* c, const void* s, int kind, const void* params, uint32_t* out_id) {
  if (!c || !s || !params) return MOX_ERR_INVALID;
  int m = addMaterial(c, kind, params);
  if (m < 0) return MOX_ERR_INVALID;
  ::SphereParams sp; memcpy(&sp, s, sizeof sp);
  c->spheres.push_back(sp);
  c->prims.push_back({PT_SPHERE, (int)c->spheres.size() - 1, 0, m, true});
  if (out_id) *out_id = (uint32_t)c->prims.size() - 1;
  c->built = false;
  return MOX_OK;
}
int ref_add_quad(ref_ctx* c, const void* q, int kind, const void* params, uint32_t* out_id) {
  if (!c || !q || !params) return MOX_ERR_INVALID;
  int m = addMaterial(c, kind, params);
  if (m < 0) return MOX_ERR_INVALID;
  ::QuadParams qp; memcpy(&qp, q, sizeof qp);
  c->quads.push_back(qp);
  c->prims.push_back({PT_QUAD, (int)c->quads.size() - 1, 0, m, true});
  if (out_id) *out_id = (uint32_t)c->prims.size() - 1;
  c->built = false;
  return MOX_OK;
}
// One OBJ shape with its own buffers, exactly as MinimalOptiX.cpp:392-441 uploads it.
int ref_add_mesh(ref_ctx* c, const float* v, size_t nv, const float* n, size_t nn, const float* uv, size_t nt,
                 const int32_t* vIdx, const int32_t* nIdx, const int32_t* tIdx, size_t nFaces, int kind,
                 const void* params, uint32_t* out_first) {
  if (!c || !params || (nFaces && (!v || !vIdx))) return MOX_ERR_INVALID;
  int m = addMaterial(c, kind, params);
  if (m < 0) return MOX_ERR_INVALID;
  Mesh mesh;
  for (size_t i = 0; i < nv; ++i) mesh.v.push_back(a3(v + 3 * i));
  if (n && nIdx) for (size_t i = 0; i < nn; ++i) mesh.n.push_back(a3(n + 3 * i));
  if (uv && tIdx) for (size_t i = 0; i < nt; ++i) mesh.uv.push_back(float2{uv[2 * i], uv[2 * i + 1]});
  for (size_t f = 0; f < nFaces; ++f) {
    mesh.vi.push_back(int3{vIdx[3 * f], vIdx[3 * f + 1], vIdx[3 * f + 2]});
    mesh.ni.push_back(nIdx ? int3{nIdx[3 * f], nIdx[3 * f + 1], nIdx[3 * f + 2]} : int3{0, 0, 0});
    mesh.ti.push_back(tIdx ? int3{tIdx[3 * f], tIdx[3 * f + 1], tIdx[3 * f + 2]} : int3{0, 0, 0});
  }
  c->meshes.push_back(std::move(mesh));
  if (out_first) *out_first = (uint32_t)c->prims.size();
  for (size_t f = 0; f < nFaces; ++f) c->prims.push_back({PT_TRI, (int)c->meshes.size() - 1, (int)f, m, true});
  c->built = false;
  return MOX_OK;
}
int ref_update_sphere(ref_ctx* c, uint32_t prim, const void* s) {
  if (!c || !s || prim >= c->prims.size() || c->prims[prim].type != PT_SPHERE) return MOX_ERR_INVALID;
  memcpy(&c->spheres[c->prims[prim].geom], s, sizeof(::SphereParams));
  c->built = false;
  return MOX_OK;
}
int ref_set_lights(ref_ctx* c, const void* l, size_t n) {
  if (!c || (n && !l)) return MOX_ERR_INVALID;
  c->lights.resize(n);
  if (n) memcpy(c->lights.data(), l, n * sizeof(::LightParams));
  return MOX_OK;
}
int ref_clear_scene(ref_ctx* c) {
  if (!c) return MOX_ERR_INVALID;
  c->prims.clear(); c->spheres.clear(); c->quads.clear(); c->meshes.clear(); c->mats.clear(); c->lights.clear();
  c->textures.clear(); c->built = false;
  return MOX_OK;
}
// "validate()": run the bounding-box programs; invalid boxes drop out of the acceleration structure.
int ref_build_accel(ref_ctx* c, uint32_t, float* out_ms) {
  if (!c) return MOX_ERR_INVALID;
  for (auto& p : c->prims) {
    float b[6];
    primBounds(*c, p, b);
    p.valid = p.type == PT_SPHERE ? true : (b[0] <= b[3] && b[1] <= b[4] && b[2] <= b[5]);
  }
  if (out_ms) *out_ms = 0.f;
  c->built = true;
  return MOX_OK;
}
int ref_launch(ref_ctx* c, int32_t seed) {
  if (!c) return MOX_ERR_INVALID;
  if (!c->built || !c->W) { c->err = "launch before build_accel / set_globals"; return MOX_ERR_STATE; }
  launchOnce(c, seed);
  return MOX_OK;
}
// The product's fixed seed schedule (the reference draws std::random_device seeds, utils_host.cpp:118-122).
int ref_render(ref_ctx* c, uint32_t spp, uint32_t seed) {
  if (!c) return MOX_ERR_INVALID;
  if (!c->built || !c->W) { c->err = "render before build_accel / set_globals"; return MOX_ERR_STATE; }
  for (uint32_t i = 0; i < spp; ++i) launchOnce(c, (int32_t)tea<16>((uint32_t)c->launches, seed));
  return MOX_OK;
}
int ref_read_accum(ref_ctx* c, float* dst) {
  if (!c || !dst) return MOX_ERR_INVALID;
  memcpy(dst, c->accu.data(), c->accu.size() * sizeof(float3));
  return MOX_OK;
}
int ref_map_accum(ref_ctx* c, const float** out) { if (!c || !out) return MOX_ERR_INVALID; *out = (const float*)c->accu.data(); return MOX_OK; }
int ref_unmap_accum(ref_ctx* c) { return c ? MOX_OK : MOX_ERR_INVALID; }
int ref_device_count(const ref_ctx* c) { return c ? 1 : 0; }
int ref_read_accum_begin(ref_ctx* c) { return c ? MOX_OK : MOX_ERR_INVALID; }
int ref_read_accum_end(ref_ctx* c, const float** out) { return ref_map_accum(c, out); }
int ref_clear_accum(ref_ctx* c) {
  if (!c) return MOX_ERR_INVALID;
  std::fill(c->accu.begin(), c->accu.end(), float3{0, 0, 0});
  c->launches = 0; c->primary = 0; c->bounce = 0; c->shadow = 0;
  return MOX_OK;
}
int ref_set_accum(ref_ctx* c, const float* src, uint64_t launches) {
  if (!c || !src || c->accu.empty()) return MOX_ERR_INVALID;
  memcpy(c->accu.data(), src, c->accu.size() * sizeof(float3));
  c->launches = launches;
  return MOX_OK;
}
int ref_owned_pixels(ref_ctx* c, uint32_t rank, uint64_t* n) { if (!c || !n || rank) return MOX_ERR_INVALID; *n = (uint64_t)c->W * c->H; return MOX_OK; }
int ref_pack_owned(ref_ctx* c, void*) { if (c) c->err = "single-device"; return MOX_ERR_INVALID; }
int ref_unpack_owned(ref_ctx* c, uint32_t, const void*) { if (c) c->err = "single-device"; return MOX_ERR_INVALID; }
int ref_get_stats(ref_ctx* c, mox_stats* s) {
  if (!c || !s) return MOX_ERR_INVALID;
  *s = mox_stats{};
  s->rays_primary = c->primary; s->rays_bounce = c->bounce; s->rays_shadow = c->shadow; s->launches = c->launches;
  s->n_prims = (uint32_t)c->prims.size(); s->n_spheres = (uint32_t)c->spheres.size(); s->n_quads = (uint32_t)c->quads.size();
  s->n_triangles = s->n_prims - s->n_spheres - s->n_quads; s->n_lights = (uint32_t)c->lights.size();
  return MOX_OK;
}
int ref_get_device_stats(ref_ctx* c, int index, mox_stats* s) { return index == 0 ? ref_get_stats(c, s) : MOX_ERR_INVALID; }
// Exception.cu:10-12 for one pixel.
int ref_exception(ref_ctx* c, uint32_t x, uint32_t y) {
  if (!c || x >= c->W || y >= c->H) return MOX_ERR_INVALID;
  bindContext(c);
  ref_exc::launchIdx = uint2{x, y};
  ref_exc::exception();
  return MOX_OK;
}

// Closest hit of a radiance ray as the intersection programs report it.  hits: n x 4 words
// (t, prim id or -1, beta, gamma) — beta/gamma are recomputed with the SDK triangle test for
// the winning triangle (they are locals of meshIntersect).  attrs (may be NULL): n x 15 floats
// geoNormal, shadingNormal, frontHitPoint, backHitPoint, texcoord.
int ref_trace_closest_attrs(ref_ctx* c, const float* rays, size_t n, void* hits, float* attrs) {
  if (!c || (n && (!rays || !hits))) return MOX_ERR_INVALID;
  if (!c->built) { c->err = "trace before build_accel"; return MOX_ERR_STATE; }
  bindContext(c);
  for (size_t i = 0; i < n; ++i) {
    const float* r = rays + 8 * i;
    Ray ray(a3(r), a3(r + 4), 0, r[3], r[7]);
    TraceState st;
    st.c = c; st.ray = ray; st.shadowRay = false; st.payload = nullptr; st.tmaxCur = ray.tmax;
    g_trace = &st;
    for (int id = 0; id < (int)c->prims.size(); ++id) {
      if (!c->prims[id].valid) continue;
      st.curPrim = id;
      intersectPrim(*c, id, ray);
    }
    g_trace = nullptr;
    float* ho = (float*)hits + 4 * i;
    int32_t* hi = (int32_t*)hits + 4 * i;
    ho[0] = ray.tmax; hi[1] = -1; ho[2] = ho[3] = 0.f;
    if (st.haveHit) {
      ho[0] = st.hitT; hi[1] = st.hitPrim;
      const Prim& p = c->prims[st.hitPrim];
      if (p.type == PT_TRI) {
        const Mesh& m = c->meshes[p.geom];
        int3 vi = m.vi[p.face];
        float3 nn; float t, be, ga;
        optix::intersect_triangle(ray, m.v[vi.x], m.v[vi.y], m.v[vi.z], nn, t, be, ga);
        ho[2] = be; ho[3] = ga;
      }
      if (attrs) {
        float* a = attrs + 15 * i;
        s3(a, st.hitAttr.geoNormal); s3(a + 3, st.hitAttr.shadingNormal); s3(a + 6, st.hitAttr.front);
        s3(a + 9, st.hitAttr.back); s3(a + 12, st.hitAttr.texcoord);
      }
    } else if (attrs) {
      for (int k = 0; k < 15; ++k) attrs[15 * i + k] = 0.f;
    }
  }
  return MOX_OK;
}
int ref_trace_closest(ref_ctx* c, const float* rays, size_t n, void* hits) { return ref_trace_closest_attrs(c, rays, n, hits, nullptr); }
// Shadow rays as Material.cu:185-191 traces them: attenuation starts at 1, any-hit programs run.
int ref_trace_shadow(ref_ctx* c, const float* rays, size_t n, float* out) {
  if (!c || (n && (!rays || !out))) return MOX_ERR_INVALID;
  if (!c->built) { c->err = "trace before build_accel"; return MOX_ERR_STATE; }
  bindContext(c);
  for (size_t i = 0; i < n; ++i) {
    const float* r = rays + 8 * i;
    Ray ray(a3(r), a3(r + 4), 1, r[3], r[7]);
    ::Payload p{};
    p.depth = 2; p.attenuation = float3{1.f, 1.f, 1.f};
    refshim::trace(ray, &p);
    s3(out + 3 * i, p.attenuation);
  }
  return MOX_OK;
}
// Bounding-box program output for primitive `prim` (6 floats: min, max as written by the program).
int ref_prim_bounds(ref_ctx* c, uint32_t prim, float out[6]) {
  if (!c || prim >= c->prims.size()) return MOX_ERR_INVALID;
  primBounds(*c, c->prims[prim], out);
  return MOX_OK;
}

// ---- the device helpers one by one (utils_device.h, disney.h)
uint32_t ref_tea16(uint32_t a, uint32_t b) { return tea<16>(a, b); }
uint32_t ref_lcg(int32_t* seed) { int s = *seed; uint32_t v = lcg(s); *seed = s; return v; }
float ref_rand(int32_t* seed) { int s = *seed; float v = rand(s); *seed = s; return v; }
void ref_rand_in_unit_sphere(int32_t* seed, float out[3]) { int s = *seed; s3(out, randInUnitSphere(s)); *seed = s; }
void ref_rand_in_unit_disk(int32_t* seed, float out[3]) { int s = *seed; s3(out, randInUnitDisk(s)); *seed = s; }
int32_t ref_fork_seed(int32_t parentSeed, int32_t parentDepth) {
  ::Payload p{}; p.randSeed = parentSeed; p.depth = parentDepth;
  return folkPayload(p).randSeed;
}
float ref_fresnel(float ci, float ct, float ior) { return fresnel(ci, ct, ior); }
void ref_offset(const float hit[3], const float n[3], float out[3]) { s3(out, offset(a3(hit), a3(n))); }
void ref_refine_hitpoint(const float hit[3], const float dir[3], const float n[3], const float p[3], float back[3],
                         float front[3]) {
  float3 b, f;
  refineHitpoint(a3(hit), a3(dir), a3(n), a3(p), b, f);
  s3(back, b); s3(front, f);
}
float ref_gtr1(float ndh, float a) { return GTR1(ndh, a); }
float ref_gtr2(float ndh, float a) { return GTR2(ndh, a); }
float ref_gtr2_aniso(float ndh, float hx, float hy, float ax, float ay) { return GTR2Aniso(ndh, hx, hy, ax, ay); }
float ref_schlick_fresnel(float u) { return schlickFresnel(u); }
float ref_smith_ggx(float ndv, float a) { return smithGGgx(ndv, a); }
float ref_smith_ggx_aniso(float ndv, float vx, float vy, float ax, float ay) { return smithGGgxAniso(ndv, vx, vy, ax, ay); }
float ref_power_heuristic(float a, float b) { return powerHeuristic(a, b); }
void ref_srgb2lin(const float v[3], float out[3]) { s3(out, srgb2lin(a3(v))); }
void ref_disney_eval(const void* mp, const float bc[3], const float N[3], const float L[3], const float V[3],
                     const float H[3], float out[3]) {
  ::DisneyParams d; memcpy(&d, mp, sizeof d);
  float3 b = a3(bc), n = a3(N), l = a3(L), v = a3(V), h = a3(H);
  s3(out, disneyEval(d, b, n, l, v, h));
}
float ref_disney_pdf(const void* mp, const float N[3], const float L[3], const float V[3], const float H[3]) {
  ::DisneyParams d; memcpy(&d, mp, sizeof d);
  float3 n = a3(N), l = a3(L), v = a3(V), h = a3(H);
  return disneyPdf(d, n, l, v, h);
}
void ref_disney_sample(int32_t* seed, const void* mp, const float N[3], const float V[3], float L[3], float H[3]) {
  ::DisneyParams d; memcpy(&d, mp, sizeof d);
  int s = *seed;
  float3 n = a3(N), v = a3(V), l{0, 0, 0}, h{0, 0, 0};
  disneySample(s, d, n, l, v, h);
  *seed = s; s3(L, l); s3(H, h);
}
// SDK helpers as the shim restates them (not reference text; exported so tests can show the two
// restatements, shim and oracle/vecmath.h, agree).
int ref_refract(const float i[3], const float n[3], float ior, float out[3]) {
  float3 r; bool ok = optix::refract(r, a3(i), a3(n), ior); s3(out, r); return ok ? 1 : 0;
}

}  // extern "C"
