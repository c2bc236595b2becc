#include "scene.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "hvec.h"
#include "utils_host.h"

using namespace moxh;

namespace {

const int kMaxLine = 2048;  // the reference reads with fgets(line, 2048)

// If `line` is <blanks><key><rest>, return rest; else nullptr.  This is what a scanf format
// " key ..." accepts: leading blanks optional, the key matched literally, anything after.
const char* afterKey(const char* line, const char* key) {
  while (isspace((unsigned char)*line)) ++line;
  size_t n = strlen(key);
  return strncmp(line, key, n) == 0 ? line + n : nullptr;
}

// Parse `count` floats from rest; stores as many as parse (scanf assigns left to right and
// stops at the first failure).
void readFloats(const char* rest, float* dst, int count) {
  for (int i = 0; i < count; ++i) {
    char* end = nullptr;
    float f = strtof(rest, &end);
    if (end == rest) return;
    dst[i] = f;
    rest = end;
  }
}

bool readWord(const char* rest, std::string& out) {
  while (isspace((unsigned char)*rest)) ++rest;
  size_t n = 0;
  while (rest[n] && !isspace((unsigned char)rest[n])) ++n;
  if (n == 0) return false;
  out.assign(rest, n);
  return true;
}

void floatKey(const char* line, const char* key, float* dst, int count) {
  if (const char* r = afterKey(line, key)) readFloats(r, dst, count);
}

struct LineReader {
  FILE* f;
  char buf[kMaxLine];
  bool next() { return fgets(buf, kMaxLine, f) != nullptr; }
};

}  // namespace

Scene::Scene(const char* fileName) {
  FILE* file = fopen(fileName, "r");
  if (!file) {
    ok = false;
    error = std::string("Couldn't open ") + fileName + " for reading.";
    return;
  }
  std::map<std::string, DisneyParams> materialMap;
  std::map<std::string, std::string> textureMap;
  LineReader in{file, {0}};

  while (in.next()) {
    char* line = in.buf;
    if (line[0] == '#') continue;  // comments only at column 0

    // material NAME { key value ... }
    std::string name;
    if (const char* r = afterKey(line, "material")) {
      if (readWord(r, name)) {
        DisneyParams m;
        initDisneyParams(m);
        std::string texName;
        while (in.next()) {
          if (strchr(line, '}')) break;
          if (const char* q = afterKey(line, "name")) readWord(q, name);
          floatKey(line, "color", &m.color.x, 3);
          if (const char* q = afterKey(line, "albedoTex")) readWord(q, texName);
          floatKey(line, "emission", &m.emission.x, 3);
          floatKey(line, "metallic", &m.metallic, 1);
          floatKey(line, "subsurface", &m.subsurface, 1);
          floatKey(line, "specular", &m.specular, 1);          // "specularTint ..." fails the float parse: no effect
          floatKey(line, "specularTint", &m.specularTint, 1);
          floatKey(line, "roughness", &m.roughness, 1);
          floatKey(line, "anisotropic", &m.anisotropic, 1);
          floatKey(line, "sheen", &m.sheen, 1);
          floatKey(line, "sheenTint", &m.sheenTint, 1);
          floatKey(line, "clearcoat", &m.clearcoat, 1);
          floatKey(line, "clearcoatGloss", &m.clearcoatGloss, 1);
          if (const char* q = afterKey(line, "brdf")) {
            char* end = nullptr;
            long v = strtol(q, &end, 0);  // %i: decimal, 0x.., 0..
            if (end != q) m.brdfType = (BrdfType)v;
          }
        }
        m.albedoID = MOX_TEXTURE_ID_NULL;
        materialMap[name] = m;
        textureMap[name] = texName;
      }
    }

    // light { ... } — detected by substring, as the reference does (scene.cpp:59)
    if (strstr(line, "light")) {
      LightParams light;
      memset(&light, 0, sizeof light);
      float3 v1 = mk3(0.f), v2 = mk3(0.f);
      std::string type = "None";
      while (in.next()) {
        if (strchr(line, '}')) break;
        floatKey(line, "position", &light.position.x, 3);
        floatKey(line, "emission", &light.emission.x, 3);
        floatKey(line, "normal", &light.normal.x, 3);
        floatKey(line, "radius", &light.radius, 1);
        floatKey(line, "v1", &v1.x, 3);
        floatKey(line, "v2", &v2.x, 3);
        if (const char* q = afterKey(line, "type")) readWord(q, type);
      }
      if (type == "Quad") {
        light.shape = QUAD;
        light.u = v1 - light.position;
        light.v = v2 - light.position;
        light.area = length(cross(light.u, light.v));
        light.normal = normalize(cross(light.u, light.v));
      } else if (type == "Sphere") {
        light.shape = SPHERE;
        light.normal = normalize(light.normal);
        light.area = 4.0f * 3.14159265358979323846f * light.radius * light.radius;
      }
      lights.push_back(light);
    }

    if (strstr(line, "properties")) {
      while (in.next()) {
        if (strchr(line, '}')) break;
        if (const char* q = afterKey(line, "width")) { char* e; long v = strtol(q, &e, 0); if (e != q) width = (int)v; }
        if (const char* q = afterKey(line, "height")) { char* e; long v = strtol(q, &e, 0); if (e != q) height = (int)v; }
      }
    }

    if (strstr(line, "mesh")) {
      while (in.next()) {
        if (strchr(line, '}')) break;
        std::string word;
        if (const char* q = afterKey(line, "file")) {
          if (readWord(q, word)) meshNames.push_back(word);
        }
        if (const char* q = afterKey(line, "material")) {
          if (readWord(q, word)) {
            auto it = materialMap.find(word);
            if (it != materialMap.end()) {
              materials.push_back(it->second);
              textures.push_back(textureMap[word]);
            } else {
              ok = false;
              error = "Could not find material " + word;
            }
          }
        }
      }
    }
  }
  fclose(file);
  if (ok && (meshNames.size() != materials.size())) {
    ok = false;
    error = "mesh block without file or material";
  }
}
