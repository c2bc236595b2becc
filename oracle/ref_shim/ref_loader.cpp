// TEST INFRASTRUCTURE.  C wrapper around the REFERENCE's own loader code — scene.cpp and
// tiny_obj_loader.h are compiled from /root/reference (never copied) — used in this container
// to generate golden fixtures (scripts/make_loader_golden.py) and to differential-test our
// from-scratch Scene parser and OBJ reader.  The render half of the reference needs OptiX and
// cannot be built.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "scene.h"  // the reference's (include path -I$(REF))
#define TINYOBJLOADER_IMPLEMENTATION
#include "tiny_obj_loader.h"  // the reference's

// utils_host.cpp cannot be compiled (libav / NVRTC); the reference's Scene only needs the Disney
// defaults from it (utils_host.cpp:101-116), supplied here.
void initDisneyParams(DisneyParams& d) {
  d.color = optix::make_float3(1.0f, 1.0f, 1.0f);
  d.emission = optix::make_float3(0.0f);
  d.metallic = 0.0f; d.subsurface = 0.0f; d.specular = 0.5f; d.roughness = 0.5f; d.specularTint = 0.0f;
  d.anisotropic = 0.0f; d.sheen = 0.0f; d.sheenTint = 0.5f; d.clearcoat = 0.0f; d.clearcoatGloss = 1.0f;
  d.brdfType = NORMAL; d.albedoID = RT_TEXTURE_ID_NULL;
}

static uint64_t fnv1a(const void* data, size_t n) {
  const uint8_t* p = (const uint8_t*)data;
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

extern "C" {

struct RefScene { Scene* s; };

void* ref_scene_open(const char* path) { return new RefScene{new Scene(path)}; }
void ref_scene_close(void* h) { RefScene* r = (RefScene*)h; delete r->s; delete r; }
int ref_scene_counts(void* h, int* meshes, int* materials, int* lights, int* width, int* height) {
  Scene* s = ((RefScene*)h)->s;
  *meshes = (int)s->meshNames.size(); *materials = (int)s->materials.size(); *lights = (int)s->lights.size();
  *width = s->width; *height = s->height;
  return 0;
}
const char* ref_scene_mesh_name(void* h, int i) { return ((RefScene*)h)->s->meshNames[i].c_str(); }
const char* ref_scene_texture(void* h, int i) { return ((RefScene*)h)->s->textures[i].c_str(); }
void ref_scene_material(void* h, int i, void* out72) { memcpy(out72, &((RefScene*)h)->s->materials[i], sizeof(DisneyParams)); }
void ref_scene_light(void* h, int i, void* out72) { memcpy(out72, &((RefScene*)h)->s->lights[i], sizeof(LightParams)); }

// Load an OBJ with the reference's tinyobj; report per-shape counts and FNV-1a hashes of the
// attribute arrays and of the (v, vn, vt) index triples, in the layout moxh_scene_mesh_hash uses.
int ref_obj_load(const char* path, int* n_shapes, uint64_t* n_vertices, uint64_t* n_normals, uint64_t* n_texcoords,
                 uint64_t attr_hash[3], uint64_t* faces /*[max_shapes]*/, uint64_t* index_hash /*[max_shapes]*/, int max_shapes) {
  tinyobj::attrib_t attrib;
  std::vector<tinyobj::shape_t> shapes;
  std::vector<tinyobj::material_t> materials;
  std::string warn, err;
  bool ret = tinyobj::LoadObj(&attrib, &shapes, &materials, &warn, &err, path);
  if (!err.empty() || !ret) return -1;
  *n_shapes = (int)shapes.size();
  *n_vertices = attrib.vertices.size() / 3; *n_normals = attrib.normals.size() / 3; *n_texcoords = attrib.texcoords.size() / 2;
  attr_hash[0] = fnv1a(attrib.vertices.data(), attrib.vertices.size() * 4);
  attr_hash[1] = fnv1a(attrib.normals.data(), attrib.normals.size() * 4);
  attr_hash[2] = fnv1a(attrib.texcoords.data(), attrib.texcoords.size() * 4);
  for (int s = 0; s < (int)shapes.size() && s < max_shapes; ++s) {
    std::vector<int32_t> idx;
    for (auto& i : shapes[s].mesh.indices) { idx.push_back(i.vertex_index); idx.push_back(i.normal_index); idx.push_back(i.texcoord_index); }
    faces[s] = shapes[s].mesh.num_face_vertices.size();
    index_hash[s] = fnv1a(idx.data(), idx.size() * 4);
  }
  return 0;
}

int ref_parse_double(const char* text, double* out) { return tinyobj::tryParseDouble(text, text + strlen(text), out) ? 1 : 0; }

}  // extern "C"
