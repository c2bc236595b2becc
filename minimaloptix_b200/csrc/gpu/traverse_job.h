// traverse_job.h — launch descriptor of the persistent traversal kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef MOX_TRAV_TPB
#define MOX_TRAV_TPB 128   // threads per CTA of the persistent traversal kernels (64 x 18 CTAs per SM and 256 x 4 measured: see DESIGN.md)
#endif

// One batch of rays for the persistent traversal kernel.
//   rayO[id] = (origin, tmin)   rayD[id] = (direction, tmax); tmax < 0 marks an unused slot
//   queue    : optional indirection, ray id = queue[i] for i < count
//   cursor   : global fetch cursor, zeroed before the launch
//   hits     : closest hit, raw query form: hits[id] = (t, prim id as int bits or -1, beta, gamma)
//   hits2    : closest hit, render form (CLASSIFY kernels): hits2[id] = (t, prim id | class << 28, or -1)
//   shC      : any hit: shC[id].xyz *= transmittance
struct TraceJob {
  const float4* rayO;
  const float4* rayD;
  const uint32_t* queue;
  uint32_t count;
  const uint32_t* countPtr;  // when set, the ray count is read from device memory instead
  uint32_t originMod;        // when non-zero, the origin of ray id is rayO[id % originMod]
  uint32_t originMagic;      // floor(2^32 / originMod) + 1: id / originMod = umulhi(id, magic) or one less (set by launchTraverse)
  uint32_t* cursor;
  float4* hits;
  float2* hits2;
  float4* shC;
  uint32_t* counters;
  int fetchThreshold;   // refill idle lanes when fewer than this many lanes are traversing
};

