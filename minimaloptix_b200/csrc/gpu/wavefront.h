// wavefront.h — device buffers and launchers of the wavefront integrator.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "gpu_types.h"
#include "traverse_job.h"

// Material queues built by the extend kernel.
enum ShadeQueue { Q_LAMBERT = 0, Q_METAL = 1, Q_DIELECTRIC = 2, Q_DISNEY = 3, Q_COUNT = 4 };

// Device counters (uint32 words).  Words 0..7 exist once per bounce in a small ring (LaunchCtx::bc points at the
// current bounce's block), so the host never has to zero them between two kernels of the same bounce loop:
//   C_NEXT  rays spawned by this bounce's shade kernels (= the next extend launch's ray count, read on the device)
//   C_MAT0+k paths binned into material queue k by this bounce's classify kernel
//   C_SHQ   shadow rays queued by this bounce's Disney kernel
// Words 8.. are per batch.
enum CounterSlot { C_NEXT = 0, C_MAT0 = 1 /* ..4 */, C_SHQ = 5 /* shadow queue length */, C_BOUNCE_WORDS = 8,
                   C_NONFINITE = 8, C_SHADOW = 9, C_NODEVIS_LO = 10, C_NODEVIS_HI = 11,
                   C_PRIMTEST_LO = 12, C_PRIMTEST_HI = 13, C_CURSOR = 14, C_CURSOR_SHADOW = 15 /* the shadow launch may overlap the next extend launch */,
                   C_NODEVIS_SH_LO = 16, C_PRIMTEST_SH_LO = 18, C_SH_BLOCKED = 20, C_SH_TINTED = 21, C_WORDS = 24 };
constexpr int BOUNCE_RING = 4;   // per-bounce counter blocks kept alive at once

struct PathBuffers {
  size_t capacity = 0;      // paths
  size_t shadowSlots = 0;   // capacity * nLights
  float4 *rayO = nullptr, *rayD = nullptr, *thr = nullptr, *rad = nullptr;
  float2* hit = nullptr;    // (t, primitive id | shade class << 28 | MOX_HIT_MISS | MOX_HIT_DEAD)
  int* state = nullptr;
  uint32_t *qCur = nullptr, *qNext = nullptr;
  uint32_t* qBuf[3] = {nullptr, nullptr, nullptr};   // queue storage; qCur / qNext point into it
  uint32_t* kBuf[2] = {nullptr, nullptr};            // sort keys of qNext (ray reordering) + scratch
  uint32_t* qKey = nullptr;                          // keys written next to qNext (null: reordering off)
  uint32_t* sortScratch = nullptr;
  uint32_t* qMat[Q_COUNT] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t* qDisneyAlt = nullptr;   // second Disney queue: bounce b's k_apply may still read its queue while bounce b+1 is classified
  float4 *shO = nullptr, *shD = nullptr, *shC = nullptr;  // shadow rays: origin per Disney path, (dir, tmax) and contribution per slot
  uint32_t* shQueue = nullptr;                            // slots that need a shadow ray
  float4* disneyRec = nullptr;                            // MOX_DISNEY_REC_F4 words per Disney hit, word-major (word * capacity + hit)
  uint32_t* counters = nullptr;   // C_WORDS per-batch words (0..7 unused) followed by BOUNCE_RING blocks of C_BOUNCE_WORDS
  int32_t* seeds = nullptr;       // launch seed per sample of the batch
  size_t seedCap = 0;
};

#define MOX_DISNEY_REC_F4 7

struct LaunchCtx {
  SceneView scene;
  RenderParams rp;
  PathBuffers pb;
  uint32_t* bc;              // this bounce's counter block (C_NEXT, C_MAT0.., C_SHQ)
  const uint32_t* ownedPix;  // device: the pixels this slice renders
  uint32_t nOwned;
  float* accu;               // device W*H*3
  bool countTraversal;
  bool disneySplit;          // Disney NORMAL as two kernels (light sampling, then BSDF sampling) instead of one
  bool brdfFast;             // hardware reciprocal / square-root approximations inside BRDF values (shading.cuh::bdiv)
  float3 sceneLo, sceneInvExt;  // ray-reordering key: 7-bit cell of the origin inside the scene box
  uint32_t sortShift = 0;       // the key's low bits dropped: 0 = 24-bit key (3 sort passes), 16 = octant + 5 cell bits (1 pass)
  cudaStream_t stream;
};

void launchGenerate(const LaunchCtx& c, uint32_t nSamples);
// countPtr (may be null): exact ray count in device memory; `count` is then an upper bound used to size the grid.
void launchExtend(const LaunchCtx& c, const uint32_t* queue, uint32_t count, const uint32_t* countPtr, uint32_t depth);
void launchClassify(const LaunchCtx& c, const uint32_t* queue, uint32_t count, const uint32_t* countPtr, uint32_t depth);
void launchShade(const LaunchCtx& c, int kind, uint32_t count, uint32_t depth);
// The shadow traversal and k_apply may run on their own stream, overlapping the next bounce's extend launch.
void launchShadow(const LaunchCtx& c, uint32_t disneyCount, cudaStream_t stream);
void launchApply(const LaunchCtx& c, uint32_t disneyCount, cudaStream_t stream);
void launchAccumulate(const LaunchCtx& c, uint32_t nSamples);
// Persistent traversal over one batch of rays (closest hit or shadow transmittance).
void launchTraverse(const SceneView& s, const TraceJob& job, bool anyHit, bool count, cudaStream_t stream, bool classify = false);
void launchSplitRays(const float4* rays, float4* o, float4* d, size_t n, cudaStream_t stream);
void launchBuildShadeRecords(const TriIdx* tris, const float* verts, const float* normals, const float* uvs, uint32_t n, float4* out,
                             cudaStream_t stream);
void launchFillOnes(float4* p, size_t n, cudaStream_t stream);
void launchCopyRgb(const float4* src, float* dst, size_t n, cudaStream_t stream);
void launchPackOwned(const float* accu, const uint32_t* ownedPix, uint32_t nOwned, float* dst, cudaStream_t stream);
void launchUnpackOwned(float* accu, const uint32_t* ownedPix, uint32_t nOwned, const float* src, cudaStream_t stream);
// dst[pix] = accu[pix] for the owned pixels; dst may live on another GPU (peer stores over NVLink).
void launchPushOwned(const float* accu, const uint32_t* ownedPix, uint32_t nOwned, float* dstFull, cudaStream_t stream);
