// obj_loader.h — a from-scratch Wavefront OBJ reader exposing the subset of the
// tiny_obj_loader v1.4 API the reference calls (MinimalOptiX.cpp:380-441):
//   tinyobj::attrib_t / shape_t / mesh_t / index_t / material_t and
//   tinyobj::LoadObj(&attrib, &shapes, &materials, &warn, &err, filename).
// Behaviour that the render path depends on is reproduced, not the whole library:
//   * number parsing follows tinyobj's own digit-accumulating parser (tiny_obj_loader.h:567-680),
//     not strtod, so vertex bits are identical to what the reference uploads;
//   * indices: 1-based, negative = relative (tiny_obj_loader.h:501-522, 820-872);
//   * polygons are ear-clipped in the dominant-axis projection (tiny_obj_loader.h:1107-1310),
//     triangles pass through in file order;
//   * `g` / `o` start a new shape; `usemtl`, `mtllib`, `s`, `t`, `l` do not change geometry.
//     OBJ materials are never read by the reference (MinimalOptiX.cpp:380-389), so
//     `materials` is always returned empty.
// A differential test against the reference's header compiled in oracle/_ref pins this.
#pragma once
#include <string>
#include <vector>

namespace tinyobj {
typedef float real_t;

struct index_t { int vertex_index; int normal_index; int texcoord_index; };

struct mesh_t {
  std::vector<index_t> indices;
  std::vector<unsigned char> num_face_vertices;
  std::vector<int> material_ids;
};

struct shape_t { std::string name; mesh_t mesh; };

struct attrib_t {
  std::vector<real_t> vertices;   // xyz
  std::vector<real_t> normals;    // xyz
  std::vector<real_t> texcoords;  // uv
};

struct material_t { std::string name; };

bool LoadObj(attrib_t* attrib, std::vector<shape_t>* shapes, std::vector<material_t>* materials,
             std::string* warn, std::string* err, const char* filename, const char* mtl_basedir = nullptr,
             bool triangulate = true);

// Exposed for tests: tinyobj's number grammar.  Returns false when [s, s_end) is not a number.
bool tryParseDouble(const char* s, const char* s_end, double* result);
}  // namespace tinyobj
