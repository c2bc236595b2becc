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
};

struct BuildOutput {
  BvhNode2* nodes = nullptr;  // device, owned by the caller after a successful build
  int nNodes = 0;
  float4* packed = nullptr;   // device, 3 float4 per valid primitive in leaf order
  int nValid = 0, nInvalid = 0;
  float sceneLo[3] = {0, 0, 0}, sceneHi[3] = {0, 0, 0};
  float4 *scratchLo = nullptr, *scratchHi = nullptr;  // builder-internal
};

bool buildBvh(const BuildInput& in, BuildOutput& out, cudaStream_t stream, std::string& err);
bool radixSortPairs(uint32_t* keys, uint32_t* vals, int n, cudaStream_t stream, std::string& err);
