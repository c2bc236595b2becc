// scene_desc.h — host-side flattened scene (what MinimalOptiX::setupScene()/setupScene(name)
// push into OptiX, MinimalOptiX.cpp:154-538, gathered in one backend-independent value) and
// the builders for the scenes the reference knows plus the synthetic BASELINE configs.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "api_table.h"
#include "hvec.h"
#include "mox_structs.h"

namespace moxh {

struct MaterialBlock {
  int kind = 0;  // mox_material_kind
  union {
    LambertianParams lam;
    MetalParams met;
    GlassParams gls;
    DisneyParams dis;
    LightParams lgt;
  };
  MaterialBlock() { memset_zero(); }
  void memset_zero();
  const void* params() const { return &lam; }
};

struct MeshDesc {
  std::string name;
  std::vector<float> v, n, uv;          // whole attribute arrays of the OBJ the shape came from
  std::vector<int32_t> vi, ni, ti;      // 3 per face
  size_t faces() const { return vi.size() / 3; }
};

struct Item {  // one GeometryInstance, in insertion (= primitive id) order
  enum Type { SPHERE_ITEM, QUAD_ITEM, MESH_ITEM } type;
  SphereParams sphere;
  QuadParams quad;
  int mesh = -1;  // index into SceneDesc::meshes
  int texture = -1;  // index into SceneDesc::textures (Disney albedo), -1 = none
  MaterialBlock mat;
};

struct TextureDesc { std::string name; int w = 0, h = 0; std::vector<float> texels; };  // RGBA float, row 0 = bottom

struct CameraDesc {
  float3 lookFrom, lookAt, up;
  float vFoV = 45.f, aperture = 0.f, focus = 1.f;
};

struct SceneDesc {
  std::string name;
  uint32_t defaultWidth = 1920, defaultHeight = 1080;  // MinimalOptiX.h:82-83
  float bg[3] = {0, 0, 0};
  CameraDesc camera;
  std::vector<Item> items;
  std::vector<MeshDesc> meshes;
  std::vector<TextureDesc> textures;  // Item::texture indexes this; uploaded once each (the reference caches by file name)
  std::vector<LightParams> lights;  // context["lights"] (Disney NEE)
  Aabb aabb;                        // over referenced mesh vertices (MinimalOptiX.cpp:430-433)
  std::vector<std::string> warnings;
  uint64_t nTriangles = 0, nVertices = 0;

  void addSphere(const SphereParams& s, const MaterialBlock& m);
  void addQuad(const float3& anchor, const float3& v1, const float3& v2, const MaterialBlock& m);
  // Validates every index against the attribute arrays first; false + err on a bad one.
  bool addMesh(MeshDesc&& mesh, const MaterialBlock& m, std::string& err);
  void addMesh(MeshDesc&& mesh, const MaterialBlock& m) { std::string e; addMesh(std::move(mesh), m, e); }  // generated meshes
  CamParams camParams(uint32_t width, uint32_t height) const;
};

MaterialBlock lambert(float r, float g, float b);
MaterialBlock metal(float r, float g, float b, float fuzz);
MaterialBlock glass(float r, float g, float b, float ior);
MaterialBlock disney(const DisneyParams& d);
MaterialBlock lightMat(float r, float g, float b);

// --- builders.  Each returns false and sets err on failure.
// SCENE_SPHERES (MinimalOptiX.cpp:156-257): 3 spheres, floor quad, quad light; aperture 0.5
// (lens) or 0 (pinhole).
bool buildSpheres(SceneDesc& out, bool pinhole, uint32_t width, uint32_t height);
// "Random spheres" = setUpVideo(256) at frame 0 (MinimalOptiX.cpp:607-759), BASELINE config 2.
bool buildRandomSpheres(SceneDesc& out, int nSpheres, uint32_t seed);
// setupScene("<name>") (MinimalOptiX.cpp:359-538) + the per-scene camera table (:258-353).
bool loadSceneFile(SceneDesc& out, const std::string& sceneDir, const std::string& name, std::string& err);
// Synthetic ~targetTris-triangle interior (BASELINE config 4; the reference's dining room /
// bedroom assets are not shipped).
bool buildInterior(SceneDesc& out, uint64_t targetTris, uint32_t seed);
// Random triangle soup in the unit cube (BASELINE config 5).
bool buildSoup(SceneDesc& out, uint64_t nTris, uint64_t seed);

// Bouncing-ball animation of the sphere items (MinimalOptiX::move / animate, MinimalOptiX.cpp:562-590):
// gravity 4000, restitution 0.9, floor plane y = -0.5.  `time` is the frame step (the reference uses 0.002).
void animateSpheres(SceneDesc& s, float time);
// Re-send every sphere item through update_sphere (primitive ids follow item order).
bool applySpheres(const SceneDesc& s, const MoxApi& api, mox_ctx* ctx, std::string& err);

// Push a SceneDesc through the C ABI: set_globals, set_camera, add_* in item order, set_lights.
// Reference defaults: eps 0.001, minIntensity 0.001, absorb 0, bad 1 (MinimalOptiX.h:85-89,
// MinimalOptiX.cpp:136-151).
bool uploadScene(const SceneDesc& s, const MoxApi& api, mox_ctx* ctx, uint32_t width, uint32_t height,
                 uint32_t maxDepth, std::string& err);

}  // namespace moxh
