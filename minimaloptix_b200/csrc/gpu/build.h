// build.h — interface between the C ABI layer and the BVH builder.
#pragma once
#include <cstring>
#include <string>
#include <algorithm>
#include "gpu_types.h"

struct BuildInput {
  int nPrims = 0;
  const PrimDesc* prims = nullptr;     // device
  const TriIdx* tris = nullptr;        // device
  const float* verts = nullptr;        // device
  const Analytic* analytic = nullptr;  // device
  cudaEvent_t evStart = nullptr, evStop = nullptr;  // recorded around the build kernels when set
  bool usePloc = true;   // false: Karras radix tree (fastest build, lower quality)
  int plocRadius = 16;
};

// Scratch of the PLOC hierarchy builder (ploc.cu), allocated before the timed build.
struct PlocScratch {
  uint32_t* cid[2] = {nullptr, nullptr};
  float4 *cLo[2] = {nullptr, nullptr}, *cHi[2] = {nullptr, nullptr};
  uint32_t* nn = nullptr;
  float4 *nodeLo = nullptr, *nodeHi = nullptr;
  uint2* children = nullptr;
  uint32_t *parent = nullptr, *size = nullptr, *leafPos = nullptr, *orderedIds = nullptr;
  unsigned long long* tileSums = nullptr;
  unsigned long long* hostTotal = nullptr;  // pinned
};
bool plocAlloc(PlocScratch& s, int n, std::string& err);
void plocFree(PlocScratch& s);
bool plocBuild(PlocScratch& s, int n, const uint32_t* sortedIds, const float4* primLo, const float4* primHi, int radius,
               BvhNode2* outNodes, float rootLo[3], float rootHi[3], int* maxDepthOut, cudaStream_t stream, std::string& err);

struct BuildOutput {
  BvhNode2* nodes = nullptr;  // device, owned by the caller after a successful build
  int nNodes = 0;
  float4* packed = nullptr;   // device, 3 float4 per valid primitive in leaf order
  int nValid = 0, nInvalid = 0;
  int iterations = 0;
  bool usedPloc = false;  // false: Karras radix tree (requested, or PLOC fallback because of depth)
  int maxDepth = 0;   // deepest leaf (PLOC); the traversal stack holds MOX_STACK entries
  float sceneLo[3] = {0, 0, 0}, sceneHi[3] = {0, 0, 0};
  float4 *scratchLo = nullptr, *scratchHi = nullptr;  // builder-internal
};

bool buildBvh(const BuildInput& in, BuildOutput& out, cudaStream_t stream, std::string& err);
bool radixSortPairs(uint32_t* keys, uint32_t* vals, int n, cudaStream_t stream, std::string& err);

// Allocation-free sorter used by the render loop (see bvh_build.cu).
size_t radixSortScratchBytes(size_t maxN);
void radixSortAsync(uint32_t* keysA, uint32_t* valsA, uint32_t* keysB, uint32_t* valsB, int n, int passes, uint32_t* scratch,
                    cudaStream_t stream);
