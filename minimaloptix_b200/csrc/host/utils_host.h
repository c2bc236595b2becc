// utils_host.h — host parameter builders with the reference's names, argument meaning and
// results (reference: MinimalOptiX/utils_host.h:23-32, utils_host.cpp:67-122).  The NVRTC and
// FFmpeg helpers of that file are out of scope (our kernels are compiled ahead of time).
#pragma once
#include <cstdint>
#include "mox_structs.h"

// plane = (normalize(v2 x v1), n.anchor); v1, v2 stored divided by their squared length.
void setQuadParams(const mox_float3& anchor, const mox_float3& v1, const mox_float3& v2, QuadParams& quadParams);

// "Ray Tracing in One Weekend" thin-lens camera; aperture 0 == pinhole.  vFoV in degrees.
void setCamParams(const mox_float3& lookFrom, const mox_float3& lookAt, const mox_float3& up, float vFoV,
                  float aspect, float aperture, float focus, CamParams& camParams);

// Disney defaults: colour 1, specular .5, roughness .5, sheenTint .5, clearcoatGloss 1, rest 0.
void initDisneyParams(DisneyParams& disneyParams);

// The reference draws launch seeds from std::random_device (utils_host.cpp:118-122), which is
// not reproducible.  Ours is a fixed schedule: seed of launch k = (int) tea<16>(k, userSeed).
int32_t launchSeed(uint32_t launchIndex, uint32_t userSeed);
uint32_t tea16(uint32_t v0, uint32_t v1);
