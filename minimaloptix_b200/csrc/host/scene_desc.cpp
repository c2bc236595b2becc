#include "scene_desc.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <random>

#include "image_io.h"
#include "obj_loader.h"
#include "scene.h"
#include "utils_host.h"

namespace moxh {

void MaterialBlock::memset_zero() { memset(&lgt, 0, sizeof lgt); }

MaterialBlock lambert(float r, float g, float b) { MaterialBlock m; m.kind = MOX_MAT_LAMBERTIAN; m.lam.albedo = mk3(r, g, b); return m; }
MaterialBlock metal(float r, float g, float b, float fuzz) { MaterialBlock m; m.kind = MOX_MAT_METAL; m.met.albedo = mk3(r, g, b); m.met.fuzz = fuzz; return m; }
MaterialBlock glass(float r, float g, float b, float ior) { MaterialBlock m; m.kind = MOX_MAT_GLASS; m.gls.albedo = mk3(r, g, b); m.gls.refIdx = ior; return m; }
MaterialBlock disney(const DisneyParams& d) { MaterialBlock m; m.kind = MOX_MAT_DISNEY; m.dis = d; return m; }
MaterialBlock lightMat(float r, float g, float b) { MaterialBlock m; m.kind = MOX_MAT_LIGHT; m.lgt.emission = mk3(r, g, b); return m; }

void SceneDesc::addSphere(const SphereParams& s, const MaterialBlock& m) {
  Item it; it.type = Item::SPHERE_ITEM; it.sphere = s; it.mat = m; memset(&it.quad, 0, sizeof it.quad);
  items.push_back(it);
}
void SceneDesc::addQuad(const float3& anchor, const float3& v1, const float3& v2, const MaterialBlock& m) {
  Item it; it.type = Item::QUAD_ITEM; it.mat = m; memset(&it.sphere, 0, sizeof it.sphere);
  memset(&it.quad, 0, sizeof it.quad);
  setQuadParams(anchor, v1, v2, it.quad);
  items.push_back(it);
}
bool SceneDesc::addMesh(MeshDesc&& mesh, const MaterialBlock& m, std::string& err) {
  const int nv = (int)(mesh.v.size() / 3), nn = (int)(mesh.n.size() / 3), nt = (int)(mesh.uv.size() / 2);
  for (size_t f = 0; f < mesh.vi.size(); ++f) {
    if (mesh.vi[f] < 0 || mesh.vi[f] >= nv) { err = "face " + std::to_string(f / 3) + " names vertex " + std::to_string(mesh.vi[f]) + " of " + std::to_string(nv); return false; }
    // -1 = "no normal / texcoord" (tinyobj); anything else must exist
    if (f < mesh.ni.size() && mesh.ni[f] != -1 && (mesh.ni[f] < 0 || mesh.ni[f] >= nn)) { err = "face " + std::to_string(f / 3) + " names normal " + std::to_string(mesh.ni[f]) + " of " + std::to_string(nn); return false; }
    if (f < mesh.ti.size() && mesh.ti[f] != -1 && (mesh.ti[f] < 0 || mesh.ti[f] >= nt)) { err = "face " + std::to_string(f / 3) + " names texcoord " + std::to_string(mesh.ti[f]) + " of " + std::to_string(nt); return false; }
  }
  for (size_t f = 0; f < mesh.vi.size(); ++f) {
    int i = mesh.vi[f];
    aabb.include(mk3(mesh.v[3 * i], mesh.v[3 * i + 1], mesh.v[3 * i + 2]));
  }
  nTriangles += mesh.faces();
  nVertices += mesh.v.size() / 3;
  meshes.push_back(std::move(mesh));
  Item it; it.type = Item::MESH_ITEM; it.mesh = (int)meshes.size() - 1; it.mat = m;
  memset(&it.sphere, 0, sizeof it.sphere); memset(&it.quad, 0, sizeof it.quad);
  items.push_back(it);
  return true;
}
CamParams SceneDesc::camParams(uint32_t width, uint32_t height) const {
  CamParams c;
  memset(&c, 0, sizeof c);
  setCamParams(camera.lookFrom, camera.lookAt, camera.up, camera.vFoV, (float)width / (float)height, camera.aperture,
               camera.focus, c);
  return c;
}

// ------------------------------------------------------------------ SCENE_SPHERES
bool buildSpheres(SceneDesc& s, bool pinhole, uint32_t, uint32_t) {
  s = SceneDesc();
  s.name = pinhole ? "spheres_pinhole" : "spheres_lens";
  s.bg[0] = s.bg[1] = s.bg[2] = 0.5f;  // MinimalOptiX.cpp:165
  SphereParams mid{0.5f, mk3(0.f, 0.f, -1.f), mk3(0.f)};
  SphereParams right{0.5f, mk3(1.f, 0.f, -1.f), mk3(0.f, 0.5f, 0.f)};
  SphereParams left{0.5f, mk3(-1.f, 0.f, -1.f), mk3(0.f, -1.5f, 0.f)};
  // group child order = primitive ids (MinimalOptiX.cpp:242): mid, floor, light, right, left
  s.addSphere(mid, lambert(0.1f, 0.2f, 0.5f));
  s.addQuad(mk3(-1000.f, -0.5f, -1000.f), mk3(2000.f, 0.f, 0.f), mk3(0.f, 0.f, 2000.f), lambert(0.8f, 0.8f, 0.f));
  s.addQuad(mk3(-5.f, 5.f, 5.f), mk3(0.f, 0.f, -10.f), mk3(10.f, 0.f, 0.f), lightMat(1.f, 1.f, 1.f));
  s.addSphere(right, metal(0.8f, 0.6f, 0.2f, 0.f));
  s.addSphere(left, glass(1.f, 1.f, 1.f, 1.5f));
  s.camera.lookFrom = mk3(3.f, 3.f, 2.f);
  s.camera.lookAt = mk3(0.f, 0.f, -1.f);
  s.camera.up = mk3(0.f, 1.f, 0.f);
  s.camera.vFoV = 20.f;
  s.camera.aperture = pinhole ? 0.f : 0.5f;
  s.camera.focus = length(s.camera.lookFrom - s.camera.lookAt);
  return true;
}

// ------------------------------------------------------------------ random spheres (config 2)
namespace {
// The reference uses libstdc++/MSVC distributions over std::mt19937(42), whose outputs are
// not portable between standard libraries.  We keep the engine (its raw 32-bit stream IS
// standardised) and define the mappings explicitly.
struct Mt {
  std::mt19937 eng;
  explicit Mt(uint32_t seed) : eng(seed) {}
  float uniform() { return (float)(eng() >> 8) * (1.0f / 16777216.0f); }
  int uniformInt3() { return (int)(((uint64_t)eng() * 3u) >> 32); }
  float normal(float sigma) {  // Box–Muller, one value per two draws
    double u1 = 1.0 - (double)uniform(), u2 = (double)uniform();
    return sigma * (float)(std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2));
  }
};
}  // namespace

bool buildRandomSpheres(SceneDesc& s, int nSpheres, uint32_t seed) {
  s = SceneDesc();
  s.name = "random_spheres";
  s.bg[0] = s.bg[1] = s.bg[2] = 0.2f;  // MinimalOptiX.cpp:611
  Mt rnd(seed);
  std::vector<SphereParams> sp;
  for (int i = 0; i < 3; ++i) sp.push_back({3.0f, mk3(-10.f + 10.f * i, 2.0f, 0.f), mk3(0.f)});
  for (int i = 0; i < nSpheres; ++i) {  // :632-646
    float x, z, radius;
    do {
      x = rnd.uniform() * 30.f - 15.f;
      z = rnd.uniform() * 30.f - 15.f;
      radius = 1.0f;
      for (auto& p : sp)
        radius = std::min(radius, sqrtf((x - p.center.x) * (x - p.center.x) + (z - p.center.z) * (z - p.center.z)) - p.radius);
      radius *= 0.8f;
    } while (radius < .01f);
    float h = sqrtf(x * x + z * z);
    radius = std::min(h + .5f, radius);
    sp.push_back({radius, mk3(x, h, z), mk3(0.f)});
  }
  std::vector<MaterialBlock> mats;
  mats.push_back(lambert(0.5f, 0.8f, 0.8f));  // :661-662
  mats.push_back(glass(1.f, 1.f, 1.f, 1.5f)); // :669-670 (i == 1)
  {
    float f = std::min(0.9f, std::max(0.1f, rnd.normal(0.1f) + 0.5f));  // :664-667 (i == 2)
    mats.push_back(metal(0.9f, 0.7f, 0.7f, f));
  }
  for (size_t i = 3; i < sp.size(); ++i) {  // :685-703
    float r = 0.2f + 0.8f * rnd.uniform();
    float g = 0.2f + 0.8f * rnd.uniform();
    float b = 0.2f + 0.8f * rnd.uniform();
    int type = rnd.uniformInt3();
    if (type == 0) mats.push_back(lambert(r, g, b));
    else if (type == 1) mats.push_back(metal(r, g, b, std::min(0.9f, std::max(0.1f, rnd.normal(0.1f) + 0.5f))));
    else mats.push_back(glass(1.f, 1.f, 1.f, std::min(3.0f, std::max(1.5f, rnd.normal(0.1f) + 2.0f))));
  }
  // group order: floor, spheres, lights (:715,741-742)
  s.addQuad(mk3(-100.f, -0.5f, 100.f), mk3(0.f, 0.f, -200.f), mk3(200.f, 0.f, 0.f), lambert(0.7f, 0.9f, 0.9f));
  for (size_t i = 0; i < sp.size(); ++i) s.addSphere(sp[i], mats[i]);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      s.addQuad(mk3(-24.f + 10.f * i, 15.f, -24.f + 10.f * j), mk3(0.f, 0.f, -8.f), mk3(8.f, 0.f, 0.f), lightMat(1.f, 1.f, 1.f));
  const int nLight = 16;
  const double angle = 3.1415926 * 2 / nLight;
  for (int i = 0; i < nLight; ++i)
    s.addQuad(mk3((float)(40.0 * std::sin(i * angle)), 1.f, (float)(40.0 * std::cos(i * angle))), mk3(0.f, 4.f, 0.f),
              mk3((float)(10.0 * std::sin(i * angle + angle) - 10.0 * std::sin(i * angle)), 0.f,
                  (float)(10.0 * std::cos(i * angle + angle) - 10.0 * std::cos(i * angle))),
              lightMat(1.f, 1.f, 1.f));
  s.camera.lookFrom = mk3(0.f, 8.0f, 20.f);  // :751-754
  s.camera.lookAt = mk3(0.f);
  s.camera.up = mk3(0.f, 1.f, 0.f);
  s.camera.vFoV = 45.f;
  s.camera.aperture = .2f;
  s.camera.focus = 20.f;
  return true;
}

// ------------------------------------------------------------------ file scenes
namespace {
struct CamRule { const char* name; float bg; float from[3]; bool fromRelCenter; float at[3]; int atMode; float fov; };
// atMode 0: at = center + a*extent ; 1: at = from + a (absolute offset)
const CamRule kCamRules[] = {
    {"coffee", 0.f, {0.f, 0.22f, 0.25f}, false, {0.f, -0.01875f, -1.f}, 1, 45.f},        // MinimalOptiX.cpp:258-270
    {"bedroom", 0.f, {0.3f, 0.1f, 0.45f}, true, {0.05f, -0.1f, 0.f}, 0, 45.f},           // :271-283
    {"diningroom", 0.f, {-0.7f, 0.f, 0.f}, true, {0.f, 0.f, 0.f}, 0, 45.f},              // :284-296
    {"stormtrooper", 0.5f, {0.25f, 0.1f, 0.395f}, true, {0.25f, 0.1f, 0.f}, 0, 30.f},    // :297-309
    {"spaceship", 0.5f, {-0.03f, 0.03f, -0.03f}, true, {0.f, 0.f, 0.f}, 0, 45.f},        // :310-322
    {"cornell", 0.5f, {0.f, 0.f, -2.f}, true, {0.f, 0.f, 0.f}, 0, 39.3077f},             // :323-335
    {"hyperion", 0.5f, {-0.08f, 2.f, 0.f}, true, {0.f, 0.f, 0.f}, 0, 30.f},              // :336-353
    {"dragon", 0.5f, {0.05f, 0.3f, -0.005f}, true, {0.f, 0.f, 0.f}, 0, 30.f},            // :336-353
};
}  // namespace

bool loadSceneFile(SceneDesc& s, const std::string& sceneDir, const std::string& name, std::string& err) {
  s = SceneDesc();
  s.name = name;
  std::string folder = sceneDir;
  if (!folder.empty() && folder.back() != '/') folder += '/';
  std::string fileBase = name == "dragon" ? "hyperion" : name;  // MinimalOptiX.cpp:340
  Scene scene((folder + fileBase + ".scene").c_str());
  if (!scene.ok) { err = scene.error; return false; }
  for (size_t i = 0; i < scene.meshNames.size(); ++i) {
    tinyobj::attrib_t attrib;
    std::vector<tinyobj::shape_t> shapes;
    std::vector<tinyobj::material_t> materials;
    std::string warn, lerr;
    bool ret = tinyobj::LoadObj(&attrib, &shapes, &materials, &warn, &lerr, (folder + scene.meshNames[i]).c_str());
    if (!lerr.empty() || !ret) {
      // The reference throws here (MinimalOptiX.cpp:386-389).  The shipped coffee scene lacks
      // Mesh010.obj, so a missing mesh is a warning and the block is skipped.
      s.warnings.push_back("skipped mesh " + scene.meshNames[i] + ": " + lerr);
      continue;
    }
    int texIndex = -1;
    if (!scene.textures[i].empty()) {  // MinimalOptiX.cpp:445-479, cached by file name
      for (size_t t = 0; t < s.textures.size(); ++t) if (s.textures[t].name == scene.textures[i]) texIndex = (int)t;
      if (texIndex < 0) {
        TextureDesc td;
        td.name = scene.textures[i];
        std::string terr;
        if (readImageRgba(folder + td.name, td.w, td.h, td.texels, terr)) { s.textures.push_back(std::move(td)); texIndex = (int)s.textures.size() - 1; }
        else s.warnings.push_back("texture " + scene.textures[i] + " not loaded: " + terr);
      }
    }
    for (auto& sh : shapes) {
      MeshDesc m;
      m.name = scene.meshNames[i] + ":" + sh.name;
      m.v = attrib.vertices;
      m.n = attrib.normals;
      m.uv = attrib.texcoords;
      size_t nf = sh.mesh.num_face_vertices.size();
      m.vi.resize(3 * nf); m.ni.resize(3 * nf); m.ti.resize(3 * nf);
      for (size_t f = 0; f < 3 * nf; ++f) {
        m.vi[f] = sh.mesh.indices[f].vertex_index;
        m.ni[f] = sh.mesh.indices[f].normal_index;
        m.ti[f] = sh.mesh.indices[f].texcoord_index;
      }
      // The OBJ reader keeps tinyobj's index arithmetic (any positive or relative index is
      // accepted), so a malformed file can name a vertex that does not exist: refuse the file,
      // as mox_add_mesh would at upload.
      if (!s.addMesh(std::move(m), disney(scene.materials[i]), err)) {
        err = scene.meshNames[i] + ": " + err;
        return false;
      }
      s.items.back().texture = texIndex;
    }
  }
  for (auto& light : scene.lights) {  // MinimalOptiX.cpp:495-521
    MaterialBlock lm; lm.kind = MOX_MAT_LIGHT; lm.lgt = light;
    if (light.shape == SPHERE) {
      SphereParams p; p.radius = light.radius; p.center = light.position; p.velocity = mk3(0.f);
      s.addSphere(p, lm);
    } else {
      s.addQuad(light.position, light.u, light.v, lm);
    }
  }
  s.lights = scene.lights;
  // properties{ width W  height H } (scene.cpp:93-101): the reference parses them and then renders at
  // its compiled-in 1920x1080; here they are the scene's default image size, which the
  // CLI uses when --width/--height are not given.
  if (scene.width > 0 && scene.height > 0) { s.defaultWidth = (uint32_t)scene.width; s.defaultHeight = (uint32_t)scene.height; }

  const CamRule* rule = nullptr;
  for (auto& r : kCamRules) if (name == r.name) rule = &r;
  if (!rule) rule = &kCamRules[5];  // unknown scene names are framed like the Cornell box
  float3 c = s.aabb.valid() ? s.aabb.center() : mk3(0.f), e = s.aabb.valid() ? s.aabb.extent() : mk3(1.f);
  float3 f = mk3(rule->from[0], rule->from[1], rule->from[2]) * e;
  s.camera.lookFrom = rule->fromRelCenter ? c + f : f;
  float3 a = mk3(rule->at[0], rule->at[1], rule->at[2]);
  s.camera.lookAt = rule->atMode == 1 ? s.camera.lookFrom + a : c + a * e;
  s.camera.up = mk3(0.f, 1.f, 0.f);
  s.camera.vFoV = rule->fov;
  s.camera.aperture = 0.f;
  s.camera.focus = 1.f;
  s.bg[0] = s.bg[1] = s.bg[2] = rule->bg;
  return true;
}

// ------------------------------------------------------------------ procedural meshes
namespace {

struct MeshBuilder {
  MeshDesc m;
  int addVertex(float3 p) { m.v.push_back(p.x); m.v.push_back(p.y); m.v.push_back(p.z); return (int)(m.v.size() / 3) - 1; }
  void addTri(int a, int b, int c) { m.vi.push_back(a); m.vi.push_back(b); m.vi.push_back(c); }
  float3 vert(int i) const { return mk3(m.v[3 * i], m.v[3 * i + 1], m.v[3 * i + 2]); }
  // Area-weighted vertex normals; normal index == vertex index.
  void smoothNormals() {
    size_t nv = m.v.size() / 3;
    std::vector<float3> acc(nv, mk3(0.f));
    for (size_t f = 0; f < m.vi.size() / 3; ++f) {
      int a = m.vi[3 * f], b = m.vi[3 * f + 1], c = m.vi[3 * f + 2];
      float3 n = cross(vert(b) - vert(a), vert(c) - vert(a));
      acc[a] = acc[a] + n; acc[b] = acc[b] + n; acc[c] = acc[c] + n;
    }
    m.n.resize(3 * nv);
    for (size_t i = 0; i < nv; ++i) {
      float l = length(acc[i]);
      float3 n = l > 0 ? acc[i] / l : mk3(0.f, 1.f, 0.f);
      m.n[3 * i] = n.x; m.n[3 * i + 1] = n.y; m.n[3 * i + 2] = n.z;
    }
    m.ni = m.vi;
  }
};

// Unit icosphere, `level` subdivisions (20 * 4^level faces), CCW seen from outside.
void icosphere(MeshBuilder& b, int level) {
  const float t = (1.0f + sqrtf(5.0f)) * 0.5f;
  const float3 base[12] = {mk3(-1, t, 0), mk3(1, t, 0), mk3(-1, -t, 0), mk3(1, -t, 0), mk3(0, -1, t), mk3(0, 1, t),
                           mk3(0, -1, -t), mk3(0, 1, -t), mk3(t, 0, -1), mk3(t, 0, 1), mk3(-t, 0, -1), mk3(-t, 0, 1)};
  static const int faces[20][3] = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                                   {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8},
                                   {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
  for (auto& p : base) b.addVertex(normalize(p));
  std::vector<int> tri;
  for (auto& f : faces) { tri.push_back(f[0]); tri.push_back(f[1]); tri.push_back(f[2]); }
  for (int l = 0; l < level; ++l) {
    std::map<uint64_t, int> mid;
    auto midpoint = [&](int a, int c) {
      uint64_t key = ((uint64_t)std::min(a, c) << 32) | (uint32_t)std::max(a, c);
      auto it = mid.find(key);
      if (it != mid.end()) return it->second;
      int idx = b.addVertex(normalize((b.vert(a) + b.vert(c)) * 0.5f));
      mid[key] = idx;
      return idx;
    };
    std::vector<int> next;
    next.reserve(tri.size() * 4);
    for (size_t f = 0; f < tri.size(); f += 3) {
      int a = tri[f], c = tri[f + 1], d = tri[f + 2];
      int ac = midpoint(a, c), cd = midpoint(c, d), da = midpoint(d, a);
      int q[12] = {a, ac, da, c, cd, ac, d, da, cd, ac, cd, da};
      next.insert(next.end(), q, q + 12);
    }
    tri.swap(next);
  }
  for (size_t f = 0; f < tri.size(); f += 3) b.addTri(tri[f], tri[f + 1], tri[f + 2]);
}

// Lumpy sphere: radial displacement by a few sine lobes.
MeshDesc blob(int level, float3 center, float radius, float lump, uint32_t seed) {
  MeshBuilder b;
  icosphere(b, level);
  std::mt19937 eng(seed);
  float ph[9];
  for (float& p : ph) p = (float)(eng() >> 8) * (6.2831853f / 16777216.0f);
  size_t nv = b.m.v.size() / 3;
  for (size_t i = 0; i < nv; ++i) {
    float3 p = b.vert((int)i);
    float d = sinf(3.f * p.x + ph[0]) * sinf(4.f * p.y + ph[1]) * sinf(5.f * p.z + ph[2]) +
              0.5f * sinf(9.f * p.x + ph[3]) * sinf(7.f * p.y + ph[4]) * sinf(8.f * p.z + ph[5]) +
              0.25f * sinf(17.f * p.x + ph[6]) * sinf(19.f * p.y + ph[7]) * sinf(23.f * p.z + ph[8]);
    float3 q = center + p * (radius * (1.f + lump * d));
    b.m.v[3 * i] = q.x; b.m.v[3 * i + 1] = q.y; b.m.v[3 * i + 2] = q.z;
  }
  b.smoothNormals();
  return std::move(b.m);
}

MeshDesc torus(int nu, int nv, float3 center, float R, float r, float tilt) {
  MeshBuilder b;
  float ct = cosf(tilt), st = sinf(tilt);
  for (int i = 0; i < nu; ++i)
    for (int j = 0; j < nv; ++j) {
      float u = 6.2831853f * i / nu, v = 6.2831853f * j / nv;
      float3 p = mk3((R + r * cosf(v)) * cosf(u), r * sinf(v), (R + r * cosf(v)) * sinf(u));
      float3 q = mk3(p.x, p.y * ct - p.z * st, p.y * st + p.z * ct);
      b.addVertex(center + q);
    }
  for (int i = 0; i < nu; ++i)
    for (int j = 0; j < nv; ++j) {
      int a = i * nv + j, c = ((i + 1) % nu) * nv + j, d = ((i + 1) % nu) * nv + (j + 1) % nv, e = i * nv + (j + 1) % nv;
      b.addTri(a, d, c);
      b.addTri(a, e, d);
    }
  b.smoothNormals();
  return std::move(b.m);
}

// Flat rectangle origin + s*eu + t*ev tessellated nu x nv, normal = eu x ev.
MeshDesc gridQuad(float3 origin, float3 eu, float3 ev, int nu, int nv) {
  MeshBuilder b;
  for (int j = 0; j <= nv; ++j)
    for (int i = 0; i <= nu; ++i) b.addVertex(origin + eu * ((float)i / nu) + ev * ((float)j / nv));
  for (int j = 0; j < nv; ++j)
    for (int i = 0; i < nu; ++i) {
      int a = j * (nu + 1) + i, c = a + 1, d = a + nu + 2, e = a + nu + 1;
      b.addTri(a, c, d);
      b.addTri(a, d, e);
    }
  return std::move(b.m);
}

MeshDesc boxMesh(float3 lo, float3 hi) {
  MeshBuilder b;
  for (int k = 0; k < 8; ++k) b.addVertex(mk3(k & 1 ? hi.x : lo.x, k & 2 ? hi.y : lo.y, k & 4 ? hi.z : lo.z));
  static const int q[6][4] = {{0, 2, 3, 1}, {4, 5, 7, 6}, {0, 1, 5, 4}, {2, 6, 7, 3}, {0, 4, 6, 2}, {1, 3, 7, 5}};
  for (auto& f : q) { b.addTri(f[0], f[1], f[2]); b.addTri(f[0], f[2], f[3]); }
  return std::move(b.m);
}

DisneyParams dparams(float r, float g, float b, float rough, float metallic, BrdfType type = NORMAL) {
  DisneyParams d;
  initDisneyParams(d);
  d.color = mk3(r, g, b); d.roughness = rough; d.metallic = metallic; d.brdfType = type;
  return d;
}

void addQuadLight(SceneDesc& s, float3 pos, float3 u, float3 v, float e) {
  LightParams l;
  memset(&l, 0, sizeof l);
  l.shape = QUAD; l.position = pos; l.u = u; l.v = v;
  l.emission = mk3(e);
  l.area = length(cross(u, v));
  l.normal = normalize(cross(u, v));
  s.lights.push_back(l);
}

}  // namespace

bool buildInterior(SceneDesc& s, uint64_t targetTris, uint32_t seed) {
  s = SceneDesc();
  s.name = "interior";
  s.bg[0] = s.bg[1] = s.bg[2] = 0.f;
  std::mt19937 eng(seed);
  auto U = [&]() { return (float)(eng() >> 8) * (1.0f / 16777216.0f); };
  const float X = 10.f, Y = 4.f, Z = 8.f;  // room [0,X] x [0,Y] x [0,Z]
  // Room shell, normals facing inward.
  s.addMesh(gridQuad(mk3(0, 0, 0), mk3(0, 0, Z), mk3(X, 0, 0), 32, 32), disney(dparams(0.578f, 0.578f, 0.578f, 0.3f, 0.f)));  // floor (+y)
  s.addMesh(gridQuad(mk3(0, Y, 0), mk3(X, 0, 0), mk3(0, 0, Z), 8, 8), disney(dparams(0.8f, 0.8f, 0.8f, 0.5f, 0.f)));          // ceiling (-y)
  s.addMesh(gridQuad(mk3(0, 0, 0), mk3(X, 0, 0), mk3(0, Y, 0), 8, 8), disney(dparams(0.75f, 0.75f, 0.7f, 0.5f, 0.f)));        // back z=0 (+z)
  s.addMesh(gridQuad(mk3(0, 0, Z), mk3(0, Y, 0), mk3(X, 0, 0), 8, 8), disney(dparams(0.75f, 0.75f, 0.7f, 0.5f, 0.f)));        // front z=Z (-z)
  s.addMesh(gridQuad(mk3(0, 0, 0), mk3(0, Y, 0), mk3(0, 0, Z), 8, 8), disney(dparams(0.7f, 0.25f, 0.2f, 0.5f, 0.f)));         // left x=0 (+x)
  s.addMesh(gridQuad(mk3(X, 0, 0), mk3(0, 0, Z), mk3(0, Y, 0), 8, 8), disney(dparams(0.2f, 0.3f, 0.7f, 0.5f, 0.f)));          // right x=X (-x)
  // A long table: top slab and four legs.
  s.addMesh(boxMesh(mk3(2.f, 0.95f, 2.5f), mk3(8.f, 1.05f, 5.5f)), disney(dparams(0.35f, 0.2f, 0.1f, 0.2f, 0.f)));
  for (int k = 0; k < 4; ++k) {
    float lx = k & 1 ? 7.7f : 2.1f, lz = k & 2 ? 5.2f : 2.6f;
    s.addMesh(boxMesh(mk3(lx, 0.f, lz), mk3(lx + 0.2f, 0.95f, lz + 0.2f)), disney(dparams(0.3f, 0.18f, 0.1f, 0.3f, 0.f)));
  }
  uint64_t budget = targetTris > s.nTriangles ? targetTris - s.nTriangles : 0;
  const int level = budget >= 400000 ? 5 : budget >= 60000 ? 4 : 3;
  const uint64_t blobTris = 20ull << (2 * level);
  const int tu = budget >= 400000 ? 192 : 64, tv = budget >= 400000 ? 64 : 24;
  const uint64_t torusTris = 2ull * tu * tv;
  uint64_t nBlobs = (uint64_t)((double)budget * 0.75 / (double)blobTris + 0.5);
  uint64_t nTori = budget > nBlobs * blobTris ? (budget - nBlobs * blobTris + torusTris / 2) / torusTris : 0;
  uint64_t nObj = nBlobs + nTori;
  // Objects: first fill the table top on a jittered grid, the rest stand on the floor.
  int cols = std::max(1, (int)std::ceil(std::sqrt((double)nObj * 1.6)));
  for (uint64_t k = 0; k < nObj; ++k) {
    bool isBlob = k < nBlobs;
    int gx = (int)(k % cols), gz = (int)(k / cols);
    int rows = (int)((nObj + cols - 1) / cols);
    float fx = (gx + 0.5f + 0.3f * (U() - 0.5f)) / cols, fz = (gz + 0.5f + 0.3f * (U() - 0.5f)) / std::max(1, rows);
    float px = 0.6f + fx * (X - 1.2f), pz = 0.6f + fz * (Z - 1.2f);
    bool onTable = px > 2.2f && px < 7.8f && pz > 2.7f && pz < 5.3f;
    float cell = std::min((X - 1.2f) / cols, (Z - 1.2f) / std::max(1, rows));
    float rad = 0.32f * cell * (0.8f + 0.4f * U());
    float baseY = onTable ? 1.05f : 0.f;
    float hue = U();
    int style = (int)(k % 7);
    DisneyParams d;
    if (style == 3) d = dparams(1.f, 1.f, 1.f, 0.5f, 0.f, GLASS);
    else if (style == 5) d = dparams(0.95f, 0.85f + 0.1f * hue, 0.6f, 0.05f + 0.2f * U(), 1.f);
    else d = dparams(0.2f + 0.8f * hue, 0.2f + 0.8f * U(), 0.2f + 0.8f * U(), 0.05f + 0.5f * U(), 0.f);
    if (isBlob) {
      s.addMesh(blob(level, mk3(px, baseY + rad * 1.25f, pz), rad, 0.18f, (uint32_t)eng()), disney(d));
    } else {
      float r = rad * 0.3f;
      s.addMesh(torus(tu, tv, mk3(px, baseY + rad * 0.8f + r, pz), rad * 0.8f, r, 1.2f * (U() - 0.5f)), disney(d));
    }
  }
  // Four ceiling panels facing down (u x v = -y), geometry + NEE entries (as setupScene does for
  // `light` blocks, MinimalOptiX.cpp:495-531).
  for (int k = 0; k < 4; ++k) {
    float cx = k & 1 ? 7.f : 3.f, cz = k & 2 ? 5.8f : 2.2f;
    addQuadLight(s, mk3(cx - 0.75f, Y - 0.01f, cz - 0.75f), mk3(1.5f, 0.f, 0.f), mk3(0.f, 0.f, 1.5f), 6.f);
  }
  for (auto& l : s.lights) { MaterialBlock lm; lm.kind = MOX_MAT_LIGHT; lm.lgt = l; s.addQuad(l.position, l.u, l.v, lm); }
  s.camera.lookFrom = mk3(0.9f, 2.3f, 7.4f);
  s.camera.lookAt = mk3(5.5f, 0.9f, 3.5f);
  s.camera.up = mk3(0.f, 1.f, 0.f);
  s.camera.vFoV = 45.f;
  s.camera.aperture = 0.f;
  s.camera.focus = 1.f;
  s.defaultWidth = 3840; s.defaultHeight = 2160;
  return true;
}

bool buildSoup(SceneDesc& s, uint64_t nTris, uint64_t seed) {
  s = SceneDesc();
  s.name = "soup";
  s.bg[0] = s.bg[1] = s.bg[2] = 0.5f;
  std::mt19937_64 eng(seed);
  auto U = [&]() { return (float)(eng() >> 40) * (1.0f / 16777216.0f); };
  const float sz = 0.5f * (float)std::pow((double)std::max<uint64_t>(nTris, 1), -1.0 / 3.0);
  MeshDesc m;
  m.name = "soup";
  m.v.resize(nTris * 9);
  m.vi.resize(nTris * 3);
  for (uint64_t t = 0; t < nTris; ++t) {
    float cx = U(), cy = U(), cz = U();
    for (int k = 0; k < 3; ++k) {
      m.v[9 * t + 3 * k + 0] = cx + (2.f * U() - 1.f) * sz;
      m.v[9 * t + 3 * k + 1] = cy + (2.f * U() - 1.f) * sz;
      m.v[9 * t + 3 * k + 2] = cz + (2.f * U() - 1.f) * sz;
      m.vi[3 * t + k] = (int32_t)(3 * t + k);
    }
  }
  s.addMesh(std::move(m), lambert(0.7f, 0.7f, 0.7f));
  s.camera.lookFrom = mk3(0.5f, 0.5f, 3.0f);
  s.camera.lookAt = mk3(0.5f, 0.5f, 0.5f);
  s.camera.up = mk3(0.f, 1.f, 0.f);
  s.camera.vFoV = 30.f;
  s.camera.aperture = 0.f;
  s.camera.focus = 1.f;
  return true;
}

// ------------------------------------------------------------------ animation
namespace {
const float kGravity = 4000.f, kRestitution = 0.9f, kFloorY = -0.5f;

// One sphere, `time` seconds: free fall until the floor plane, then a damped bounce.
void moveSphere(SphereParams& p, float time, int depth = 0) {
  float fall = p.velocity.y * time + time * time * kGravity / 2.0f;  // distance travelled downwards
  float room = p.center.y - p.radius - kFloorY;                      // height above the floor
  if (fall < room) {
    p.center.x += p.velocity.x * time;
    p.center.z += p.velocity.z * time;
    p.center.y -= fall;
    p.velocity.y += kGravity * time;
    return;
  }
  float vend = sqrtf(p.velocity.y * p.velocity.y + 2.0f * kGravity * room);  // speed at impact
  float t = (vend - p.velocity.y) / kGravity;                                // time to impact
  if (t < 1e-6f || depth > 64) {  // resting on the floor
    p.velocity.y = 0.f;
    p.center.y = kFloorY + p.radius;
    return;
  }
  p.center.x += p.velocity.x * t;
  p.center.z += p.velocity.z * t;
  p.center.y = kFloorY + p.radius;
  p.velocity.x *= kRestitution;
  p.velocity.y = -vend * kRestitution;  // velocity.y is positive downwards
  moveSphere(p, time - t, depth + 1);
}
}  // namespace

void animateSpheres(SceneDesc& s, float time) {
  for (Item& it : s.items)
    if (it.type == Item::SPHERE_ITEM && it.mat.kind != MOX_MAT_LIGHT) moveSphere(it.sphere, time);
}

bool applySpheres(const SceneDesc& s, const MoxApi& api, mox_ctx* ctx, std::string& err) {
  uint32_t prim = 0;
  for (const Item& it : s.items) {
    if (it.type == Item::SPHERE_ITEM && api.update_sphere(ctx, prim, &it.sphere)) { err = std::string("update_sphere: ") + api.last_error(ctx); return false; }
    prim += it.type == Item::MESH_ITEM ? (uint32_t)s.meshes[it.mesh].faces() : 1u;
  }
  return true;
}

// ------------------------------------------------------------------ upload through the C ABI
bool uploadScene(const SceneDesc& s, const MoxApi& api, mox_ctx* ctx, uint32_t width, uint32_t height, uint32_t maxDepth,
                 std::string& err) {
  auto fail = [&](const char* what) { err = std::string(what) + ": " + api.last_error(ctx); return false; };
  const float absorb[3] = {0.f, 0.f, 0.f}, bad[3] = {1.f, 1.f, 1.f};
  if (api.clear_scene(ctx)) return fail("clear_scene");
  if (api.set_globals(ctx, width, height, maxDepth, 0.001f, 0.001f, absorb, bad, s.bg)) return fail("set_globals");
  CamParams cam = s.camParams(width, height);
  if (api.set_camera(ctx, &cam)) return fail("set_camera");
  std::vector<int> texIds(s.textures.size(), 0);
  for (size_t t = 0; t < s.textures.size(); ++t)
    if (api.add_texture_rgba32f(ctx, s.textures[t].texels.data(), s.textures[t].w, s.textures[t].h, &texIds[t])) return fail("add_texture");
  for (const Item& itRef : s.items) {
    Item it = itRef;
    if (it.texture >= 0 && it.mat.kind == MOX_MAT_DISNEY) it.mat.dis.albedoID = texIds[it.texture];
    int rc = 0;
    if (it.type == Item::SPHERE_ITEM) rc = api.add_sphere(ctx, &it.sphere, it.mat.kind, it.mat.params(), nullptr);
    else if (it.type == Item::QUAD_ITEM) rc = api.add_quad(ctx, &it.quad, it.mat.kind, it.mat.params(), nullptr);
    else {
      const MeshDesc& m = s.meshes[it.mesh];
      rc = api.add_mesh(ctx, m.v.data(), m.v.size() / 3, m.n.empty() ? nullptr : m.n.data(), m.n.size() / 3,
                        m.uv.empty() ? nullptr : m.uv.data(), m.uv.size() / 2, m.vi.data(),
                        m.ni.empty() ? nullptr : m.ni.data(), m.ti.empty() ? nullptr : m.ti.data(), m.faces(), it.mat.kind,
                        it.mat.params(), nullptr);
    }
    if (rc) return fail("add geometry");
  }
  if (api.set_lights(ctx, s.lights.data(), s.lights.size())) return fail("set_lights");
  return true;
}

}  // namespace moxh
