/* mox_structs.h — parameter blocks shared by the host driver, the C ABI, the CUDA
 * kernels and the CPU oracle.
 *
 * These are the POD blocks MinimalOptiX hands to OptiX through setUserData()
 * (reference: MinimalOptiX/Structures.h:5-80).  Names, field order, sizes and
 * offsets are kept so a MinimalOptiX host can pass its structs straight through
 * the C ABI in mox.h; the vector types are plain C structs that are layout-
 * compatible with CUDA's float3 / float4 (the optix::float3 the reference uses
 * is that same CUDA type).
 *
 * Sizes (checked at compile time below):
 *   Payload 32, CamParams 76, SphereParams 28, QuadParams 64 (align 16),
 *   LambertianParams 12, MetalParams 16, GlassParams 16, DisneyParams 72,
 *   LightParams 72.
 */
#ifndef MOX_STRUCTS_H
#define MOX_STRUCTS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define MOX_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define MOX_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

typedef struct mox_float2 { float x, y; } mox_float2;
typedef struct mox_float3 { float x, y, z; } mox_float3;
typedef struct __attribute__((aligned(16))) mox_float4 { float x, y, z, w; } mox_float4;

/* Texture id 0 means "no texture" (RT_TEXTURE_ID_NULL in OptiX 5). */
#define MOX_TEXTURE_ID_NULL 0

/* Per-ray payload of the reference's recursive integrator (Structures.h:5-10).
 * Our wavefront path state carries the same information in SoA form; the struct
 * is kept for the oracle and for ABI completeness. */
typedef struct Payload {
  mox_float3 color;
  int depth;
  int randSeed;
  mox_float3 attenuation; /* shadow rays only */
} Payload;

/* Thin-lens camera (pinhole == lensRadius 0), built by setCamParams(). */
typedef struct CamParams {
  mox_float3 origin;
  mox_float3 horizontal;
  mox_float3 vertical;
  mox_float3 scrLowerLeftCorner;
  mox_float3 u;
  mox_float3 v;
  float lensRadius;
} CamParams;

typedef struct SphereParams {
  float radius;
  mox_float3 center;
  mox_float3 velocity; /* animation only; ignored by the render path */
} SphereParams;

/* plane = (unit normal, normal.anchor); v1, v2 are pre-divided by |v|^2
 * (setQuadParams, utils_host.cpp:67-75). */
typedef struct QuadParams {
  mox_float4 plane;
  mox_float3 v1;
  mox_float3 v2;
  mox_float3 anchor;
} QuadParams;

typedef struct LambertianParams { mox_float3 albedo; } LambertianParams;
typedef struct MetalParams { mox_float3 albedo; float fuzz; } MetalParams;
typedef struct GlassParams { mox_float3 albedo; float refIdx; } GlassParams;

typedef enum BrdfType { NORMAL = 0, GLASS = 1 } BrdfType;

typedef struct DisneyParams {
  int albedoID;
  mox_float3 color;
  mox_float3 emission;
  float metallic;
  float subsurface;
  float specular;
  float roughness;
  float specularTint;
  float anisotropic;
  float sheen;
  float sheenTint;
  float clearcoat;
  float clearcoatGloss;
  BrdfType brdfType;
} DisneyParams;

typedef enum LightShape { SPHERE = 0, QUAD = 1 } LightShape;

typedef struct LightParams {
  mox_float3 position;
  mox_float3 normal;
  mox_float3 emission;
  mox_float3 u; /* quad edges, not normalised */
  mox_float3 v;
  float area;
  float radius;
  LightShape shape;
} LightParams;

MOX_STATIC_ASSERT(sizeof(mox_float3) == 12, "float3");
MOX_STATIC_ASSERT(sizeof(mox_float4) == 16, "float4");
MOX_STATIC_ASSERT(sizeof(Payload) == 32, "Payload");
MOX_STATIC_ASSERT(sizeof(CamParams) == 76, "CamParams");
MOX_STATIC_ASSERT(sizeof(SphereParams) == 28, "SphereParams");
MOX_STATIC_ASSERT(sizeof(QuadParams) == 64, "QuadParams");
MOX_STATIC_ASSERT(offsetof(QuadParams, v1) == 16, "QuadParams.v1");
MOX_STATIC_ASSERT(offsetof(QuadParams, v2) == 28, "QuadParams.v2");
MOX_STATIC_ASSERT(offsetof(QuadParams, anchor) == 40, "QuadParams.anchor");
MOX_STATIC_ASSERT(sizeof(LambertianParams) == 12, "LambertianParams");
MOX_STATIC_ASSERT(sizeof(MetalParams) == 16, "MetalParams");
MOX_STATIC_ASSERT(sizeof(GlassParams) == 16, "GlassParams");
MOX_STATIC_ASSERT(sizeof(DisneyParams) == 72, "DisneyParams");
MOX_STATIC_ASSERT(offsetof(DisneyParams, color) == 4, "DisneyParams.color");
MOX_STATIC_ASSERT(offsetof(DisneyParams, emission) == 16, "DisneyParams.emission");
MOX_STATIC_ASSERT(offsetof(DisneyParams, metallic) == 28, "DisneyParams.metallic");
MOX_STATIC_ASSERT(offsetof(DisneyParams, roughness) == 40, "DisneyParams.roughness");
MOX_STATIC_ASSERT(offsetof(DisneyParams, brdfType) == 68, "DisneyParams.brdfType");
MOX_STATIC_ASSERT(sizeof(LightParams) == 72, "LightParams");
MOX_STATIC_ASSERT(offsetof(LightParams, u) == 36, "LightParams.u");
MOX_STATIC_ASSERT(offsetof(LightParams, area) == 60, "LightParams.area");
MOX_STATIC_ASSERT(offsetof(LightParams, shape) == 68, "LightParams.shape");

#endif /* MOX_STRUCTS_H */
