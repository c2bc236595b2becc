#include "utils_host.h"
#include "hvec.h"

using namespace moxh;

void setQuadParams(const mox_float3& anchor, const mox_float3& v1, const mox_float3& v2, QuadParams& q) {
  float3 n = normalize(cross(v2, v1));
  q.plane.x = n.x; q.plane.y = n.y; q.plane.z = n.z;
  q.plane.w = dot(n, anchor);
  q.v1 = v1 / dot(v1, v1);
  q.v2 = v2 / dot(v2, v2);
  q.anchor = anchor;
}

void setCamParams(const mox_float3& lookFrom, const mox_float3& lookAt, const mox_float3& up, float vFoV,
                  float aspect, float aperture, float focus, CamParams& cam) {
  // All in float, as the reference's float locals make it (utils_host.cpp:81-84).
  const float pi = 3.14159265358979323846f;
  float theta = vFoV * pi / 180;
  float halfHeight = tanf(theta / 2);
  float halfWidth = aspect * halfHeight;
  float3 w = normalize(lookFrom - lookAt);
  float3 u = normalize(cross(up, w));
  float3 v = cross(w, u);
  cam.origin = lookFrom;
  cam.scrLowerLeftCorner = lookFrom - focus * halfWidth * u - focus * halfHeight * v - focus * w;
  cam.horizontal = 2 * focus * halfWidth * u;
  cam.vertical = 2 * focus * halfHeight * v;
  cam.u = u;
  cam.v = v;
  cam.lensRadius = aperture / 2;
}

void initDisneyParams(DisneyParams& d) {
  d.albedoID = MOX_TEXTURE_ID_NULL;
  d.color = mk3(1.0f);
  d.emission = mk3(0.0f);
  d.metallic = 0.0f;
  d.subsurface = 0.0f;
  d.specular = 0.5f;
  d.roughness = 0.5f;
  d.specularTint = 0.0f;
  d.anisotropic = 0.0f;
  d.sheen = 0.0f;
  d.sheenTint = 0.5f;
  d.clearcoat = 0.0f;
  d.clearcoatGloss = 1.0f;
  d.brdfType = NORMAL;
}

uint32_t tea16(uint32_t v0, uint32_t v1) {
  uint32_t s0 = 0;
  for (int n = 0; n < 16; n++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}

int32_t launchSeed(uint32_t launchIndex, uint32_t userSeed) { return (int32_t)tea16(launchIndex, userSeed); }
