// wavefront.h — device buffers and launchers of the wavefront integrator.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "gpu_types.h"
#include "traverse_job.h"

// Material queues built by the extend kernel.
enum ShadeQueue { Q_LAMBERT = 0, Q_METAL = 1, Q_DIELECTRIC = 2, Q_DISNEY = 3, Q_COUNT = 4 };

// Device counters (uint32 words).
enum CounterSlot { C_NEXT = 0, C_MAT0 = 1 /* ..4 */, C_SHQ = 5 /* shadow queue length */, C_NONFINITE = 8, C_SHADOW = 9, C_NODEVIS_LO = 10, C_NODEVIS_HI = 11,
                   C_PRIMTEST_LO = 12, C_PRIMTEST_HI = 13, C_CURSOR = 14, C_NODEVIS_SH_LO = 16, C_PRIMTEST_SH_LO = 18, C_WORDS = 20 };

struct PathBuffers {
  size_t capacity = 0;      // paths
  size_t shadowSlots = 0;   // capacity * nLights
  float4 *rayO = nullptr, *rayD = nullptr, *hit = nullptr, *thr = nullptr, *rad = nullptr;
  int* state = nullptr;
  uint32_t *qCur = nullptr, *qNext = nullptr;
  uint32_t* qBuf[3] = {nullptr, nullptr, nullptr};   // queue storage; qCur / qNext point into it
  uint32_t* kBuf[2] = {nullptr, nullptr};            // sort keys of qNext (ray reordering) + scratch
  uint32_t* qKey = nullptr;                          // keys written next to qNext (null: reordering off)
  uint32_t* sortScratch = nullptr;
  uint32_t* qMat[Q_COUNT] = {nullptr, nullptr, nullptr, nullptr};
  float4 *shO = nullptr, *shD = nullptr, *shC = nullptr;  // shadow rays: origin per Disney path, (dir, tmax) and contribution per slot
  uint32_t* shQueue = nullptr;                            // slots that need a shadow ray
  uint32_t* counters = nullptr;   // C_WORDS
  int32_t* seeds = nullptr;       // launch seed per sample of the batch
  size_t seedCap = 0;
};

struct LaunchCtx {
  SceneView scene;
  RenderParams rp;
  PathBuffers pb;
  const uint32_t* ownedPix;  // device
  uint32_t nOwned;
  float* accu;               // device W*H*3
  bool countTraversal;
  float3 sceneLo, sceneInvExt;  // ray-reordering key: 7-bit cell of the origin inside the scene box
  cudaStream_t stream;
};

void launchGenerate(const LaunchCtx& c, uint32_t nSamples);
void launchExtend(const LaunchCtx& c, const uint32_t* queue, uint32_t count, uint32_t depth);
void launchShade(const LaunchCtx& c, int kind, uint32_t count, uint32_t depth);
void launchLogic(const LaunchCtx& c, const uint32_t* queue, uint32_t count, uint32_t depth);
void launchShadow(const LaunchCtx& c, uint32_t disneyCount);
void launchApply(const LaunchCtx& c, uint32_t disneyCount);
void launchAccumulate(const LaunchCtx& c, uint32_t nSamples);
// Persistent traversal over one batch of rays (closest hit or shadow transmittance).
void launchTraverse(const SceneView& s, const TraceJob& job, bool anyHit, bool count, cudaStream_t stream);
void launchSplitRays(const float4* rays, float4* o, float4* d, size_t n, cudaStream_t stream);
void launchBuildShadeRecords(const TriIdx* tris, const float* verts, const float* normals, const float* uvs, uint32_t n, float4* out,
                             cudaStream_t stream);
void launchFillOnes(float4* p, size_t n, cudaStream_t stream);
void launchCopyRgb(const float4* src, float* dst, size_t n, cudaStream_t stream);
void launchPackOwned(const float* accu, const uint32_t* ownedPix, uint32_t nOwned, float* dst, cudaStream_t stream);
void launchUnpackOwned(float* accu, const uint32_t* ownedPix, uint32_t nOwned, const float* src, cudaStream_t stream);
