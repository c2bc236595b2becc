// build.h — interface between the C ABI layer and the BVH builder.
#pragma once
#include <cstring>
#include <string>
#include <algorithm>
#include "gpu_types.h"

// Grow-only device scratch shared by successive builds (per-frame rebuilds of an animated scene
// must not pay ~25 cudaMalloc/cudaFree pairs).  take() hands out 256-byte aligned slices.
struct DeviceArena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  unsigned long long* pinned = nullptr;  // 64 bytes of pinned host memory for small read-backs
  bool reserve(size_t bytes) {
    off = 0;
    if (bytes <= cap) return true;
    if (base) cudaFree(base);
    base = nullptr; cap = 0;
    if (cudaMalloc(&base, bytes) != cudaSuccess) return false;
    cap = bytes;
    return true;
  }
  template <class T> T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    if (off + bytes > cap) return nullptr;
    T* p = (T*)(base + off);
    off += bytes;
    return p;
  }
  void release() { if (base) cudaFree(base); if (pinned) cudaFreeHost(pinned); base = nullptr; pinned = nullptr; cap = off = 0; }
};

struct BuildInput {
  int nPrims = 0;
  const PrimDesc* prims = nullptr;     // device
  const TriIdx* tris = nullptr;        // device
  const float* verts = nullptr;        // device
  const Analytic* analytic = nullptr;  // device
  const GpuMaterial* mats = nullptr;   // device: the packed records carry each primitive's shadow-ray class
  cudaEvent_t evStart = nullptr, evStop = nullptr;  // recorded around the build kernels when set
  DeviceArena* arena = nullptr;  // required
  bool usePloc = true;   // false: Karras radix tree (fastest build, lower quality)
  int plocRadius = 0;    // 0 = automatic: 32 up to 2 M primitives, 16 above.  Bench scene (1 M): 8 -> 1152, 16 -> 1234, 32 -> 1247,
                         // 48 -> 1213, 64 -> 1219, 100 -> 1248 Mrays/s, build 3.5 -> 3.8 ms (16 -> 32); 10 M soup: 16 -> 13.2 ms and
                         // 652..856 Mrays/s, 32 -> 15.2 ms and 628..832
  bool useWide = true;   // collapse the PLOC tree into the compressed 8-wide BVH
  bool watertight = false;  // packed triangle records hold the raw vertices (p0, p1, p2) for the watertight test
};

// Scratch of the PLOC hierarchy builder (ploc.cu), allocated before the timed build.
struct PlocState;
struct PlocScratch {
  uint32_t* cid[2] = {nullptr, nullptr};
  float4 *cLo[2] = {nullptr, nullptr}, *cHi[2] = {nullptr, nullptr};
  uint32_t* nn = nullptr;
  float4 *nodeLo = nullptr, *nodeHi = nullptr;
  uint2* children = nullptr;
  uint32_t *parent = nullptr, *size = nullptr, *leafPos = nullptr, *orderedIds = nullptr;
  uint32_t* firstPos = nullptr;   // per inner node (id - nLeaves): leaf-order position of its leftmost leaf
  unsigned long long* tileSums = nullptr;
  PlocState* state = nullptr;               // loop state, one slot per iteration (ploc.cu)
  unsigned long long* hostTotal = nullptr;  // pinned
};
size_t plocScratchBytes(int n);
bool plocAlloc(PlocScratch& s, int n, DeviceArena& arena, std::string& err);
bool plocBuild(PlocScratch& s, int n, const uint32_t* sortedIds, const float4* primLo, const float4* primHi, int radius,
               BvhNode2* outNodes, float rootLo[3], float rootHi[3], int* maxDepthOut, cudaStream_t stream, std::string& err);

struct BuildOutput {
  // device, owned by the caller.  On entry they may hold buffers of nodesCap / packedCap records from a
  // previous build, which are reused when large enough.
  BvhNode2* nodes = nullptr;
  int nNodes = 0;
  float4* packed = nullptr;   // 3 float4 per valid primitive in leaf order
  size_t nodesCap = 0, packedCap = 0;
  BvhNode8* nodes8 = nullptr;  // compressed wide BVH (only when built from PLOC); same reuse rule
  float4* packed8 = nullptr;   // primitives in wide-leaf order
  size_t nodes8Cap = 0, packed8Cap = 0;
  int nNodes8 = 0, wideLevels = 0;
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

// Collapse of the PLOC tree into the compressed wide BVH (bvh_wide.cu).
size_t wideScratchBytes(int n);
bool wideCollapse(const PlocScratch& s, int n, uint32_t root, DeviceArena& arena, BvhNode8* outNodes, uint32_t** orderedIds8Out,
                  int* nNodesOut, int* levelsOut, cudaStream_t stream, std::string& err);
