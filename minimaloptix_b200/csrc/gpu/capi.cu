// capi.cu — implementation of the C ABI in include/mox.h: context, scene staging and upload,
// acceleration build, the spp loop, accumulation read-back, multi-GPU pack/unpack, raw ray
// queries.  Host-side counterpart of what MinimalOptiX::setupContext / setupScene /
// renderScene do through optixpp (MinimalOptiX.cpp:130-152, 154-538, 540-560).
//
// There is no CPU path in this library: every compute entry point needs a usable sm_100 GPU.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "build.h"
#include "group.h"
#include "mox.h"
#include "rng.cuh"
#include "wavefront.h"

namespace {
std::string g_createError;

enum Stage { ST_GENERATE = 0, ST_EXTEND, ST_SHADE, ST_SHADOW, ST_ACCUMULATE, ST_COUNT };

// Event pairs around the kernels of each stage; summed after the batch's final synchronize.
struct StageTimer {
  std::vector<cudaEvent_t> pool;
  struct Span { int stage; size_t e0, e1; uint32_t depth; };
  std::vector<Span> spans;
  size_t used = 0;
  size_t grab(cudaStream_t s) {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    cudaEventRecord(pool[used], s);
    return used++;
  }
  void begin(int stage, cudaStream_t s, uint32_t depth = 0) { spans.push_back({stage, grab(s), 0, depth}); }
  void end(cudaStream_t s) { spans.back().e1 = grab(s); }
  // msDepth[stage == ST_SHADOW][d]: the two traversal launches per path depth (mox_stats.ms_extend_depth / ms_shadow_depth)
  void collect(double* ms, double (*msDepth)[MOX_STATS_DEPTHS]) {
    for (auto& sp : spans) {
      float t = 0;
      if (cudaEventElapsedTime(&t, pool[sp.e0], pool[sp.e1]) != cudaSuccess) continue;
      ms[sp.stage] += t;
      if (sp.depth) msDepth[sp.stage == ST_SHADOW ? 1 : 0][std::min<uint32_t>(sp.depth, MOX_STATS_DEPTHS - 1)] += t;
    }
    spans.clear(); used = 0;
  }
  void release() { for (auto e : pool) cudaEventDestroy(e); pool.clear(); }
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};
}  // namespace

struct mox_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evFork = nullptr;
  std::string err;

  RenderParams rp{};
  bool haveGlobals = false, haveCamera = false;
  uint32_t rank = 0, world = 1, tile = 32;

  // host staging of the scene (copied at add_* time)
  std::vector<PrimDesc> prims;
  std::vector<TriIdx> tris;
  std::vector<float> verts, normals, uvs;
  std::vector<Analytic> analytic;
  std::vector<GpuMaterial> mats;
  std::vector<LightParams> lights;
  std::vector<float4> lightN;   // normalize(light.normal) per light, computed once with the device's operation sequence (vec.cuh)
  struct HostTexture { int w, h; std::vector<float> texels; };
  std::vector<HostTexture> textures;
  std::vector<cudaArray_t> texArrays;
  std::vector<cudaTextureObject_t> texObjects;
  bool texturesDirty = false;
  uint32_t nSpheres = 0, nQuads = 0;

  // device scene
  DevBuf dPrims, dTris, dVerts, dNormals, dUvs, dAnalytic, dMats, dLights, dLightN, dShadeRec;
  bool shadeRecBuilt = false;
  DevBuf dQueryO, dQueryD, dQueryCounters;  // raw ray queries
  DevBuf dQueryRays, dQueryOut, dQueryC;    // ... and their host-buffer forms: grow-only staging, no allocation per call
  DevBuf dTexObjs;
  BvhNode2* dNodes = nullptr;
  float4* dPacked = nullptr;
  size_t nodesCap = 0, packedCap = 0;
  BvhNode8* dNodes8 = nullptr;
  float4* dPacked8 = nullptr;
  size_t nodes8Cap = 0, packed8Cap = 0;
  int nNodes8 = 0;
  bool wideBuilt = false, useWide = true;
  DeviceArena buildArena;
  int nNodes = 0, nValid = 0;
  bool built = false, lightsDirty = true;
  uint32_t accelFlags = 0;

  // image
  float* dAccu = nullptr;
  uint32_t accuW = 0, accuH = 0;
  uint32_t* dOwned = nullptr;
  uint32_t nOwned = 0;
  bool ownedDirty = true;
  std::vector<DevBuf> otherOwned;  // cached owned lists of other ranks (unpack)
  std::vector<uint64_t> ownedCount; // cached |owned pixels| per rank of the current partition

  // The owned pixels of a batch can be rendered as `nSlices` independent sub-batches ("slices"), each with its own
  // path buffers and stream: while the host reads one slice's material counts back, and while that slice's
  // persistent traversal kernel drains its last rays, the other slice's kernels keep the SMs busy.  Since the
  // shadow rays of a bounce run on their own stream (overlapShadow) a single slice already has two traversal
  // launches in flight, and one slice measures as fast at full frame (1347 vs 1350 Mrays/s), faster on an eighth
  // of the frame (14.99 vs 15.16 ms per step) and end to end (1340 vs 1262): the default is 1 (MOX_SLICES=2..4).
  struct Slice {
    PathBuffers pb;
    cudaStream_t stream = nullptr;   // slice 0 uses the context stream
    cudaStream_t shadowStream = nullptr;  // shadow traversal + k_apply of bounce b, overlapping extend + classify of bounce b+1
    cudaEvent_t evShaded = nullptr, evApplied = nullptr;
    bool applyPending = false;       // evApplied must be waited for before the next kernel that touches rad / the shadow buffers
    cudaEvent_t evReady = nullptr;   // the counters of the bounce in flight have landed in hostCnt
    uint32_t* hostCnt = nullptr;     // pinned: C_WORDS + BOUNCE_RING * C_BOUNCE_WORDS
    StageTimer timer;
    // state of the batch in flight
    LaunchCtx lc;
    uint32_t S = 0, bound = 0, depth = 0, matCount[Q_COUNT] = {0, 0, 0, 0};
    int iCur = 0, iNext = 1, iSpare = 2;
    bool done = true;
  };
  static constexpr int kMaxSlices = 4;
  Slice slices[kMaxSlices];
  int nSlices = 1;
  float* pinned = nullptr;
  size_t pinnedBytes = 0;

  // Gather / asynchronous read-back: two full-size device buffers (targets of the tile pushes, possibly of other
  // GPUs), two pinned host images, a copy stream.  gatherPeer[w]: buffer w belongs to another rank (imported).
  float* dGather[2] = {nullptr, nullptr};
  int gatherPeer[2] = {0, 0};   // 0 own, 1 imported over CUDA IPC, 2 borrowed from another context of this process
  float* hostGather[2] = {nullptr, nullptr};
  size_t gatherBytes = 0;
  cudaStream_t copyStream = nullptr;
  cudaEvent_t evPushed = nullptr, evCopied[2] = {nullptr, nullptr};
  int readCur = 0;            // buffer the next mox_read_accum_begin uses
  int readPending = -1;       // buffer of the begin that has not been ended yet

  mox_group* group = nullptr; // set on the shell handle mox_create_multi returns

  // stats
  uint64_t raysPrimary = 0, raysBounce = 0, raysShadow = 0, nonfinite = 0, launches = 0, nodeVisits = 0, primTests = 0;
  uint64_t nodeVisitsShadow = 0, primTestsShadow = 0, raysShadowTraced = 0, shadowBlocked = 0, shadowTinted = 0;
  uint64_t raysDepth[MOX_STATS_DEPTHS] = {}, shadowDepth[MOX_STATS_DEPTHS] = {};
  double msDepth[2][MOX_STATS_DEPTHS] = {};   // [0] extend, [1] shadow traversal launches per path depth
  double msRender = 0, msBuild = 0;
  double msStage[ST_COUNT] = {0, 0, 0, 0, 0};
  uint64_t extendLaunches = 0, kernelLaunches = 0;
  size_t maxBatchPaths = 32u << 20;  // paths per wavefront (~280 B each with 4 lights); measured 4 Mi -> 947, 32 Mi -> 980 Mrays/s at 4K
  bool overlapShadow = true;     // shadow rays of bounce b on a second stream, concurrent with the extend launch of bounce b+1 (MOX_OVERLAP_SHADOW=0: serial)
  bool disneySplit = false;      // Disney NORMAL shading as two kernels sharing a per-hit record (MOX_DISNEY_SPLIT=0: one kernel)
  bool brdfFast = true;          // approximate reciprocal / square root inside BRDF values (MOX_BRDF_IEEE=1: IEEE, the oracle's operations)
  bool hasDisneyNormal = false;  // some material runs the Disney NORMAL program (set by mox_build_accel)
  bool sortRays = false;         // reorder the extend queue by (direction octant, origin cell) from bounce 2 on
  int sortPasses = 3;            // 3: the whole 24-bit key; 1: octant + the 5 top bits of the cell (MOX_SORT_PASSES)
  float sceneLo[3] = {0, 0, 0}, sceneHi[3] = {1, 1, 1};
};

namespace {

int fail(mox_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_createError = msg;
  return code;
}

#define CUCK(c, x)                                                                                    \
  do {                                                                                                \
    cudaError_t e_ = (x);                                                                             \
    if (e_ != cudaSuccess) return fail((c), e_ == cudaErrorMemoryAllocation ? MOX_ERR_OOM : MOX_ERR_CUDA, \
                                       std::string(#x) + ": " + cudaGetErrorString(e_));             \
  } while (0)

int bind(mox_ctx* c) {
  CUCK(c, cudaSetDevice(c->device));
  return MOX_OK;
}

int ensure(mox_ctx* c, DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return MOX_OK;
  b.release();
  CUCK(c, cudaMalloc(&b.p, std::max<size_t>(bytes, 16)));
  b.bytes = std::max<size_t>(bytes, 16);
  return MOX_OK;
}

template <class T>
int upload(mox_ctx* c, DevBuf& b, const std::vector<T>& v) {
  int rc = ensure(c, b, v.size() * sizeof(T));
  if (rc) return rc;
  if (!v.empty()) CUCK(c, cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  return MOX_OK;
}

int addMaterial(mox_ctx* c, int kind, const void* params) {
  GpuMaterial m;
  memset(&m, 0, sizeof m);
  m.kind = kind;
  switch (kind) {
    case MOX_MAT_LAMBERTIAN: memcpy(&m.lam, params, sizeof(LambertianParams)); break;
    case MOX_MAT_METAL: memcpy(&m.met, params, sizeof(MetalParams)); break;
    case MOX_MAT_GLASS: memcpy(&m.gls, params, sizeof(GlassParams)); break;
    case MOX_MAT_DISNEY: memcpy(&m.dis, params, sizeof(DisneyParams)); break;
    case MOX_MAT_LIGHT: memcpy(&m.lgt, params, sizeof(LightParams)); break;
    default: return -1;
  }
  c->mats.push_back(m);
  return (int)c->mats.size() - 1;
}

void ownedList(uint32_t W, uint32_t H, uint32_t tile, uint32_t world, uint32_t rank, std::vector<uint32_t>& out) {
  out.clear();
  uint32_t tx = (W + tile - 1) / tile, ty = (H + tile - 1) / tile;
  for (uint32_t j = 0; j < ty; ++j)
    for (uint32_t i = 0; i < tx; ++i) {
      if ((i + j) % world != rank) continue;
      for (uint32_t y = j * tile; y < std::min(H, (j + 1) * tile); ++y)
        for (uint32_t x = i * tile; x < std::min(W, (i + 1) * tile); ++x) out.push_back(y * W + x);
    }
}

int refreshOwned(mox_ctx* c) {
  if (!c->ownedDirty) return MOX_OK;
  std::vector<uint32_t> l;
  ownedList(c->rp.W, c->rp.H, c->tile, c->world, c->rank, l);
  if (c->dOwned) cudaFree(c->dOwned);
  c->dOwned = nullptr;
  CUCK(c, cudaMalloc(&c->dOwned, std::max<size_t>(l.size(), 1) * 4));
  // Stream-ordered upload: a plain cudaMemcpy from pageable memory may return before the DMA has landed, and
  // the context stream is non-blocking, so a kernel launched next could still read stale bytes.
  if (!l.empty()) { CUCK(c, cudaMemcpyAsync(c->dOwned, l.data(), l.size() * 4, cudaMemcpyHostToDevice, c->stream)); CUCK(c, cudaStreamSynchronize(c->stream)); }
  c->nOwned = (uint32_t)l.size();
  for (auto& b : c->otherOwned) b.release();
  c->otherOwned.clear();
  c->ownedCount.clear();
  c->ownedDirty = false;
  return MOX_OK;
}

void freePaths(PathBuffers& pb) {
  cudaFree(pb.rayO); cudaFree(pb.rayD); cudaFree(pb.hit); cudaFree(pb.thr); cudaFree(pb.rad); cudaFree(pb.state);
  for (auto& q : pb.qBuf) cudaFree(q);
  for (auto& k : pb.kBuf) cudaFree(k);
  cudaFree(pb.sortScratch);
  for (auto& q : pb.qMat) cudaFree(q);
  cudaFree(pb.shO); cudaFree(pb.shD); cudaFree(pb.shC); cudaFree(pb.shQueue); cudaFree(pb.counters); cudaFree(pb.seeds); cudaFree(pb.disneyRec); cudaFree(pb.qDisneyAlt);
  pb = PathBuffers();
}

constexpr size_t kCounterWords = C_WORDS + BOUNCE_RING * C_BOUNCE_WORDS;

// Bytes of wavefront state per path (see PathBuffers): rays, hit, throughput, radiance, RNG state, three queue
// buffers, two key buffers, four material queues; per light a direction, a contribution and a queue entry; one
// shadow origin when there are lights.
size_t bytesPerPath(size_t nLights) { return 16 * 4 + 8 + 4 + 3 * 4 + 2 * 4 + 5 * 4 + (nLights ? 16 + nLights * 36 : 0) + 16 * MOX_DISNEY_REC_F4; }

int ensurePaths(mox_ctx* c, PathBuffers& pb, cudaStream_t stream, size_t paths, size_t nLights, size_t nSeeds) {
  size_t slots = paths * nLights;
  if (pb.capacity < paths || pb.shadowSlots < slots || !pb.counters || (c->disneySplit && c->hasDisneyNormal && !pb.disneyRec)) {
    cudaStreamSynchronize(stream);
    // grow geometrically so that a slowly growing sample count does not reallocate on every call
    if (pb.capacity && paths > pb.capacity) paths = std::max(paths, pb.capacity + pb.capacity / 2);
    slots = paths * nLights;
    size_t seedCap = pb.seedCap;
    int32_t* seeds = pb.seeds;
    pb.seeds = nullptr;
    freePaths(pb);
    pb.seeds = seeds; pb.seedCap = seedCap;
    CUCK(c, cudaMalloc(&pb.rayO, paths * 16)); CUCK(c, cudaMalloc(&pb.rayD, paths * 16));
    CUCK(c, cudaMalloc(&pb.hit, paths * 8)); CUCK(c, cudaMalloc(&pb.thr, paths * 16));
    CUCK(c, cudaMalloc(&pb.rad, paths * 16)); CUCK(c, cudaMalloc(&pb.state, paths * 4));
    for (auto& q : pb.qBuf) CUCK(c, cudaMalloc(&q, paths * 4));
    if (c->sortRays) {
      for (auto& k : pb.kBuf) CUCK(c, cudaMalloc(&k, paths * 4));
      CUCK(c, cudaMalloc(&pb.sortScratch, radixSortScratchBytes(paths)));
    }
    for (auto& q : pb.qMat) CUCK(c, cudaMalloc(&q, paths * 4));
    CUCK(c, cudaMalloc(&pb.qDisneyAlt, paths * 4));
    if (slots) {
      CUCK(c, cudaMalloc(&pb.shO, paths * 16)); CUCK(c, cudaMalloc(&pb.shD, slots * 16)); CUCK(c, cudaMalloc(&pb.shC, slots * 16));
      CUCK(c, cudaMalloc(&pb.shQueue, slots * 4));
    }
    if (c->disneySplit && c->hasDisneyNormal) CUCK(c, cudaMalloc(&pb.disneyRec, paths * 16 * MOX_DISNEY_REC_F4));
    CUCK(c, cudaMalloc(&pb.counters, kCounterWords * 4));
    CUCK(c, cudaMemset(pb.counters, 0, kCounterWords * 4));
    pb.capacity = paths; pb.shadowSlots = slots;
  }
  if (pb.seedCap < nSeeds) {
    cudaStreamSynchronize(stream);
    cudaFree(pb.seeds);
    pb.seeds = nullptr;
    CUCK(c, cudaMalloc(&pb.seeds, nSeeds * 4));
    pb.seedCap = nSeeds;
  }
  return MOX_OK;
}

SceneView sceneView(const mox_ctx* c) {
  SceneView s;
  s.nodes = c->dNodes;
  s.packed = c->dPacked;
  s.nodes8 = (c->wideBuilt && c->useWide) ? c->dNodes8 : nullptr;
  s.packed8 = c->dPacked8;
  s.analytic = (const Analytic*)c->dAnalytic.p;
  s.prims = (const PrimDesc*)c->dPrims.p;
  s.mats = (const GpuMaterial*)c->dMats.p;
  s.verts = (const float*)c->dVerts.p;
  s.normals = (const float*)c->dNormals.p;
  s.uvs = (const float*)c->dUvs.p;
  s.tris = (const TriIdx*)c->dTris.p;
  s.shadeRec = c->shadeRecBuilt ? (const float4*)c->dShadeRec.p : nullptr;
  s.lights = (const LightParams*)c->dLights.p;
  s.lightN = (const float4*)c->dLightN.p;
  s.textures = (const cudaTextureObject_t*)c->dTexObjs.p;
  s.nLights = (int)c->lights.size();
  s.nPrims = (int)c->prims.size();
  s.nNodes8 = (uint32_t)c->nNodes8;
  s.watertight = (c->accelFlags & MOX_ACCEL_WATERTIGHT) ? 1 : 0;
  return s;
}

void freeTextures(mox_ctx* c) {
  for (auto t : c->texObjects) cudaDestroyTextureObject(t);
  for (auto a : c->texArrays) cudaFreeArray(a);
  c->texObjects.clear(); c->texArrays.clear();
}

// createTextureSampler: REPEAT wrap, LINEAR filter, normalized coordinates, float4 texels
// (MinimalOptiX.cpp:449-474) -> CUDA texture objects; the hardware filter is used as in OptiX.
int syncTextures(mox_ctx* c) {
  if (!c->texturesDirty) return MOX_OK;
  cudaStreamSynchronize(c->stream);
  freeTextures(c);
  for (auto& t : c->textures) {
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<float4>();
    cudaArray_t arr = nullptr;
    CUCK(c, cudaMallocArray(&arr, &fmt, t.w, t.h));
    c->texArrays.push_back(arr);
    CUCK(c, cudaMemcpy2DToArray(arr, 0, 0, t.texels.data(), (size_t)t.w * 16, (size_t)t.w * 16, t.h, cudaMemcpyHostToDevice));
    cudaResourceDesc res;
    memset(&res, 0, sizeof res);
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaTextureDesc td;
    memset(&td, 0, sizeof td);
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 1;
    cudaTextureObject_t obj = 0;
    CUCK(c, cudaCreateTextureObject(&obj, &res, &td, nullptr));
    c->texObjects.push_back(obj);
  }
  int rc = upload(c, c->dTexObjs, c->texObjects);
  if (rc) return rc;
  c->texturesDirty = false;
  return MOX_OK;
}

int syncLights(mox_ctx* c) {
  if (!c->lightsDirty) return MOX_OK;
  int rc = upload(c, c->dLights, c->lights);
  if (rc) return rc;
  // normalize(lp->normal) is the same for every hit (Material.cu:183): IEEE multiply / add / sqrt / divide in the
  // order of vec.cuh's normalize give the same bits on the host as in the shade kernel
  c->lightN.resize(c->lights.size());
  for (size_t i = 0; i < c->lights.size(); ++i) {
    const float3 n = normalize(mk3(c->lights[i].normal.x, c->lights[i].normal.y, c->lights[i].normal.z));
    c->lightN[i] = make_float4(n.x, n.y, n.z, 0.f);
  }
  if ((rc = upload(c, c->dLightN, c->lightN))) return rc;
  c->lightsDirty = false;
  return MOX_OK;
}

inline uint32_t* bounceBlock(const PathBuffers& pb, uint32_t depth) { return pb.counters + C_WORDS + (depth % BOUNCE_RING) * C_BOUNCE_WORDS; }

int ensureSlice(mox_ctx* c, int k) {
  mox_ctx::Slice& sl = c->slices[k];
  if (!sl.stream) {
    if (k == 0) sl.stream = c->stream;
    else CUCK(c, cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
  }
  if (!sl.evReady) CUCK(c, cudaEventCreateWithFlags(&sl.evReady, cudaEventDisableTiming));
  if (c->overlapShadow && !sl.shadowStream) {
    CUCK(c, cudaStreamCreateWithFlags(&sl.shadowStream, cudaStreamNonBlocking));
    CUCK(c, cudaEventCreateWithFlags(&sl.evShaded, cudaEventDisableTiming));
    CUCK(c, cudaEventCreateWithFlags(&sl.evApplied, cudaEventDisableTiming));
  }
  if (!sl.hostCnt) CUCK(c, cudaMallocHost(&sl.hostCnt, kCounterWords * 4));
  return MOX_OK;
}

// Enqueue one bounce of a slice up to the point where the host needs numbers: extend (persistent traversal, ray
// count read on the device), classify, and the read-back of the counters.
int sliceEnqueueExtend(mox_ctx* c, mox_ctx::Slice& sl) {
  LaunchCtx& lc = sl.lc;
  StageTimer& tm = sl.timer;
  const uint32_t depth = sl.depth;
  lc.bc = bounceBlock(sl.pb, depth);
  lc.pb.qMat[Q_DISNEY] = (depth & 1u) ? sl.pb.qMat[Q_DISNEY] : sl.pb.qDisneyAlt;   // k_apply of bounce depth-1 may still read the other one
  // bounce 1 traces exactly `bound` camera rays; later bounces read the number of spawned rays from the previous
  // bounce's counter block, `bound` (what was shaded) only sizes the grids
  const uint32_t* countPtr = depth == 1 ? nullptr : bounceBlock(sl.pb, depth - 1) + C_NEXT;
  tm.begin(ST_EXTEND, sl.stream, depth);
  // camera rays are queued in path order (k_generate: qCur[p] = p): the traversal kernel skips the indirection
  launchExtend(lc, depth == 1 ? nullptr : lc.pb.qCur, sl.bound, countPtr, depth);
  tm.end(sl.stream);
  tm.begin(ST_SHADE, sl.stream);
  launchClassify(lc, lc.pb.qCur, sl.bound, countPtr, depth);
  tm.end(sl.stream);
  c->extendLaunches++; c->kernelLaunches += 2;
  CUCK(c, cudaMemcpyAsync(sl.hostCnt, sl.pb.counters, kCounterWords * 4, cudaMemcpyDeviceToHost, sl.stream));
  CUCK(c, cudaEventRecord(sl.evReady, sl.stream));
  return MOX_OK;
}

int sliceFinish(mox_ctx* c, mox_ctx::Slice& sl) {
  StageTimer& tm = sl.timer;
  if (sl.applyPending) { CUCK(c, cudaStreamWaitEvent(sl.stream, sl.evApplied, 0)); sl.applyPending = false; }
  tm.begin(ST_ACCUMULATE, sl.stream);
  launchAccumulate(sl.lc, sl.S);
  tm.end(sl.stream);
  c->kernelLaunches++;
  CUCK(c, cudaMemcpyAsync(sl.hostCnt, sl.pb.counters, kCounterWords * 4, cudaMemcpyDeviceToHost, sl.stream));
  CUCK(c, cudaEventRecord(sl.evReady, sl.stream));
  sl.done = true;
  return MOX_OK;
}

// The counters of the bounce in flight have arrived: shade what was binned, trace the shadow rays, and enqueue
// the next bounce — or the accumulation when nothing is left.
int sliceAdvance(mox_ctx* c, mox_ctx::Slice& sl) {
  LaunchCtx& lc = sl.lc;
  StageTimer& tm = sl.timer;
  PathBuffers& pb = sl.pb;
  const uint32_t depth = sl.depth;
  const uint32_t* hostBounce = sl.hostCnt + C_WORDS + (depth % BOUNCE_RING) * C_BOUNCE_WORDS;
  const uint32_t dIdx = std::min<uint32_t>(depth, MOX_STATS_DEPTHS - 1);
  if (depth == 1) { c->raysPrimary += sl.bound; c->raysDepth[1] += sl.bound; }
  else {  // exact number of rays this bounce traced, and of shadow rays the previous one queued
    const uint32_t* prev = sl.hostCnt + C_WORDS + ((depth - 1) % BOUNCE_RING) * C_BOUNCE_WORDS;
    c->raysBounce += prev[C_NEXT];
    c->raysShadowTraced += prev[C_SHQ];
    c->raysDepth[dIdx] += prev[C_NEXT];
    c->shadowDepth[std::min<uint32_t>(depth - 1, MOX_STATS_DEPTHS - 1)] += prev[C_SHQ];
  }
  uint32_t any = 0;
  for (int k = 0; k < Q_COUNT; ++k) { sl.matCount[k] = hostBounce[C_MAT0 + k]; any += sl.matCount[k]; }
  if (!any) return sliceFinish(c, sl);
  // the shade kernels touch the radiance and rewrite the shadow-ray buffers: the previous bounce's k_apply first
  if (sl.applyPending) { CUCK(c, cudaStreamWaitEvent(sl.stream, sl.evApplied, 0)); sl.applyPending = false; }
  tm.begin(ST_SHADE, sl.stream);
  for (int k = 0; k < Q_COUNT; ++k) { launchShade(lc, k, sl.matCount[k], depth); if (sl.matCount[k]) c->kernelLaunches += (k == Q_DISNEY && lc.disneySplit) ? 2 : 1; }
  tm.end(sl.stream);
  if (sl.matCount[Q_DISNEY] && lc.scene.nLights) {
    // Shadow rays and the NEE sum depend only on this bounce's shade kernels; the next extend launch does not
    // depend on them.  On a second stream the two persistent traversal launches overlap: each one's last CTAs
    // drain while the other's first CTAs already run.
    cudaStream_t ss = sl.stream;
    if (c->overlapShadow && sl.shadowStream) {
      ss = sl.shadowStream;
      CUCK(c, cudaEventRecord(sl.evShaded, sl.stream));
      CUCK(c, cudaStreamWaitEvent(ss, sl.evShaded, 0));
    }
    tm.begin(ST_SHADOW, ss, depth);
    launchShadow(lc, sl.matCount[Q_DISNEY], ss);
    tm.end(ss);
    tm.begin(ST_SHADE, ss);
    launchApply(lc, sl.matCount[Q_DISNEY], ss);
    tm.end(ss);
    if (ss != sl.stream) { CUCK(c, cudaEventRecord(sl.evApplied, ss)); sl.applyPending = true; }
    c->kernelLaunches += 2;
  }
  // the block bounce depth+1 will use was last written BOUNCE_RING bounces ago
  CUCK(c, cudaMemsetAsync(bounceBlock(pb, depth + 1), 0, C_BOUNCE_WORDS * 4, sl.stream));
  uint32_t bound = any;   // every shaded path spawns at most one ray
  if (c->sortRays) {
    // reordering needs the exact count on the host: one more round trip (this mode is an experiment, off by default)
    CUCK(c, cudaMemcpyAsync(sl.hostCnt, pb.counters, kCounterWords * 4, cudaMemcpyDeviceToHost, sl.stream));
    CUCK(c, cudaStreamSynchronize(sl.stream));
    bound = hostBounce[C_NEXT];
  }
  if (c->sortRays && bound > 4096) {
    // 24-bit keys -> 3 passes (or their top 8 bits -> 1): the sorted queue lands in (kBuf[1], qBuf[iSpare])
    tm.begin(ST_SHADE, sl.stream);
    radixSortAsync(pb.kBuf[0], pb.qBuf[sl.iNext], pb.kBuf[1], pb.qBuf[sl.iSpare], (int)bound, c->sortPasses, pb.sortScratch, sl.stream);
    tm.end(sl.stream);
    c->kernelLaunches += 5;
    int t = sl.iCur; sl.iCur = sl.iSpare; sl.iSpare = sl.iNext; sl.iNext = t;
  } else {
    int t = sl.iCur; sl.iCur = sl.iNext; sl.iNext = t;
  }
  lc.pb.qCur = pb.qBuf[sl.iCur]; lc.pb.qNext = pb.qBuf[sl.iNext];
  sl.bound = bound;
  sl.depth = depth + 1;
  if (!bound) return sliceFinish(c, sl);
  return sliceEnqueueExtend(c, sl);
}

int slicesFor(const mox_ctx* c, size_t paths) {
  int K = c->nSlices;
  if (paths < (size_t)K * 65536 || c->nOwned < (uint32_t)K) K = 1;   // tiny batches: the second stream buys nothing
  return K;
}

// Streams, events, pinned counters and wavefront buffers for a batch of S samples per owned pixel.  renderSeeds
// calls it before its timed region, so the first launch of a context does not time its own allocations.
int prepareBatch(mox_ctx* c, uint32_t S) {
  const size_t P = (size_t)S * c->nOwned;
  const int K = slicesFor(c, P);
  for (int k = 0; k < K; ++k) {
    int rc = ensureSlice(c, k);
    if (rc) return rc;
    const uint32_t nPix = (uint32_t)((uint64_t)c->nOwned * (k + 1) / K) - (uint32_t)((uint64_t)c->nOwned * k / K);
    if ((rc = ensurePaths(c, c->slices[k].pb, c->slices[k].stream, (size_t)S * nPix, c->lights.size(), S))) return rc;
  }
  return MOX_OK;
}

// One wavefront batch: `seeds.size()` samples of every owned pixel, rendered as nSlices interleaved sub-batches.
int renderBatch(mox_ctx* c, const std::vector<int32_t>& seeds) {
  const uint32_t S = (uint32_t)seeds.size();
  const size_t P = (size_t)S * c->nOwned;
  if (P == 0) return MOX_OK;
  if (P > 0xfffffff0ull) return fail(c, MOX_ERR_INVALID, "batch too large");
  const int K = slicesFor(c, P);
  int rc;
  // slice k renders the owned pixels [first_k, first_k+1) — all S samples of a pixel stay in one slice, in launch
  // order, so the accumulated sums do not depend on the number of slices
  uint32_t first[mox_ctx::kMaxSlices + 1];
  for (int k = 0; k <= K; ++k) first[k] = (uint32_t)((uint64_t)c->nOwned * k / K);
  // all allocations first: a failure (MOX_ERR_OOM -> the caller retries with a smaller batch) must not leave
  // work of another slice in flight
  if ((rc = prepareBatch(c, S))) return rc;
  cudaEvent_t evStart = c->evFork;
  CUCK(c, cudaEventRecord(evStart, c->stream));
  for (int k = 0; k < K; ++k) {
    mox_ctx::Slice& sl = c->slices[k];
    const uint32_t nPix = first[k + 1] - first[k];
    const size_t paths = (size_t)S * nPix;
    if (k) CUCK(c, cudaStreamWaitEvent(sl.stream, evStart, 0));   // scene uploads happened on the context stream
    PathBuffers& pb = sl.pb;
    CUCK(c, cudaMemcpyAsync(pb.seeds, seeds.data(), S * 4, cudaMemcpyHostToDevice, sl.stream));
    LaunchCtx& lc = sl.lc;
    lc.scene = sceneView(c);
    lc.rp = c->rp;
    lc.pb = pb;
    lc.ownedPix = c->dOwned + first[k];
    lc.nOwned = nPix;
    lc.accu = c->dAccu;
    lc.countTraversal = (c->accelFlags & MOX_ACCEL_COUNTERS) != 0;
    lc.disneySplit = c->disneySplit && pb.disneyRec != nullptr;
    lc.brdfFast = c->brdfFast;
    lc.stream = sl.stream;
    lc.sceneLo = make_float3(c->sceneLo[0], c->sceneLo[1], c->sceneLo[2]);
    {
      float ex = c->sceneHi[0] - c->sceneLo[0], ey = c->sceneHi[1] - c->sceneLo[1], ez = c->sceneHi[2] - c->sceneLo[2];
      lc.sceneInvExt = make_float3(ex > 0 ? 128.f / ex : 0.f, ey > 0 ? 128.f / ey : 0.f, ez > 0 ? 128.f / ez : 0.f);
    }
    sl.iCur = 0; sl.iNext = 1; sl.iSpare = 2;
    lc.pb.qCur = pb.qBuf[sl.iCur]; lc.pb.qNext = pb.qBuf[sl.iNext];
    lc.pb.qKey = c->sortRays ? pb.kBuf[0] : nullptr;
    lc.sortShift = c->sortPasses == 1 ? 16u : 0u;
    lc.bc = bounceBlock(pb, 1);
    sl.S = S; sl.bound = (uint32_t)paths; sl.depth = 1; sl.done = false; sl.applyPending = false;
    CUCK(c, cudaMemsetAsync(pb.counters, 0, kCounterWords * 4, sl.stream));
    sl.timer.begin(ST_GENERATE, sl.stream);
    launchGenerate(lc, S);
    sl.timer.end(sl.stream);
    c->kernelLaunches++;
    if (!paths) { sl.done = true; CUCK(c, cudaEventRecord(sl.evReady, sl.stream)); continue; }
    if ((rc = sliceEnqueueExtend(c, sl))) return rc;
  }
  // Round-robin over the slices: wait for one slice's counters, enqueue its next stage, move on.  While the host
  // sits in cudaEventSynchronize for slice A, slice B's kernels are already queued on the GPU.
  for (int live = K; live > 0;) {
    live = 0;
    for (int k = 0; k < K; ++k) {
      mox_ctx::Slice& sl = c->slices[k];
      if (sl.done) continue;
      CUCK(c, cudaEventSynchronize(sl.evReady));
      if ((rc = sliceAdvance(c, sl))) return rc;
      if (!sl.done) live++;
    }
  }
  // join: every slice's accumulation and final counter read-back
  for (int k = 0; k < K; ++k) {
    mox_ctx::Slice& sl = c->slices[k];
    CUCK(c, cudaEventSynchronize(sl.evReady));
    if (k) CUCK(c, cudaStreamWaitEvent(c->stream, sl.evReady, 0));
    const uint32_t* host = sl.hostCnt;
    sl.timer.collect(c->msStage, c->msDepth);
    c->nonfinite += host[C_NONFINITE];
    c->shadowBlocked += host[C_SH_BLOCKED]; c->shadowTinted += host[C_SH_TINTED];
    c->raysShadow += host[C_SHADOW];
    c->nodeVisits += ((uint64_t)host[C_NODEVIS_HI] << 32) | host[C_NODEVIS_LO];
    c->primTests += ((uint64_t)host[C_PRIMTEST_HI] << 32) | host[C_PRIMTEST_LO];
    c->nodeVisitsShadow += ((uint64_t)host[C_NODEVIS_SH_LO + 1] << 32) | host[C_NODEVIS_SH_LO];
    c->primTestsShadow += ((uint64_t)host[C_PRIMTEST_SH_LO + 1] << 32) | host[C_PRIMTEST_SH_LO];
  }
  CUCK(c, cudaGetLastError());
  c->launches += S;
  return MOX_OK;
}

// Paths per batch: the configured cap, lowered so that the wavefront state fits into 60 % of the free device
// memory for this scene's light count (36 bytes per path and light).
size_t batchCap(mox_ctx* c, size_t wanted /* samples per pixel of this call */) {
  size_t cap = c->maxBatchPaths;
  // The buffers of an earlier call already hold such a batch: nothing will be allocated, so there is nothing to
  // budget.  (cudaMemGetInfo is a slow driver call — it was the 20-70 ms hiccup in front of some steps.)
  if (c->nOwned) {
    const size_t S = std::min(std::max<size_t>(1, cap / c->nOwned), wanted), nL = c->lights.size();   // samples per batch at the full cap
    const int K = slicesFor(c, S * c->nOwned);
    bool fits = true;
    for (int k = 0; k < K && fits; ++k) {
      const PathBuffers& pb = c->slices[k].pb;
      const size_t nPix = (size_t)((uint64_t)c->nOwned * (k + 1) / K) - (size_t)((uint64_t)c->nOwned * k / K);   // as prepareBatch splits them
      fits = pb.counters && pb.capacity >= S * nPix && pb.shadowSlots >= S * nPix * nL && pb.seedCap >= S;
    }
    if (fits) return cap;
  }
  size_t freeB = 0, totalB = 0;
  if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
    size_t have = 0;
    for (auto& sl : c->slices) have += sl.pb.capacity * bytesPerPath(sl.pb.capacity ? sl.pb.shadowSlots / sl.pb.capacity : 0);
    size_t budget = (size_t)((freeB + have) * 0.6);
    cap = std::min(cap, std::max<size_t>(budget / bytesPerPath(c->lights.size()), 65536));
  }
  return cap;
}

int renderSeeds(mox_ctx* c, const std::vector<int32_t>& seeds) {
  if (!c->built) return fail(c, MOX_ERR_STATE, "launch before mox_build_accel");
  if (!c->haveGlobals) return fail(c, MOX_ERR_STATE, "launch before mox_set_globals");
  if (!c->haveCamera) return fail(c, MOX_ERR_STATE, "launch before mox_set_camera");
  int rc = bind(c);
  if (rc) return rc;
  if ((rc = refreshOwned(c))) return rc;
  if ((rc = syncLights(c))) return rc;
  if ((rc = syncTextures(c))) return rc;
  if (c->nOwned == 0) { c->launches += seeds.size(); return MOX_OK; }
  size_t perBatch = std::max<size_t>(1, batchCap(c, seeds.size()) / c->nOwned);
  rc = prepareBatch(c, (uint32_t)std::min(perBatch, seeds.size()));
  if (rc && rc != MOX_ERR_OOM) return rc;   // out of memory: the loop below retries with smaller batches
  CUCK(c, cudaEventRecord(c->ev0, c->stream));
  for (size_t i = 0; i < seeds.size();) {
    const size_t n = std::min(perBatch, seeds.size() - i);
    std::vector<int32_t> chunk(seeds.begin() + i, seeds.begin() + i + n);
    rc = renderBatch(c, chunk);
    if (rc == MOX_ERR_OOM && perBatch > 1) {
      // the allocation failed after all (fragmentation, another context on the device): free the wavefront
      // state, halve the batch and render this chunk again (nothing of it was enqueued)
      cudaGetLastError();
      for (auto& sl : c->slices) { int32_t* sd = sl.pb.seeds; sl.pb.seeds = nullptr; freePaths(sl.pb); cudaFree(sd); }
      perBatch = std::max<size_t>(1, perBatch / 2);
      continue;
    }
    if (rc) return rc;
    i += n;
  }
  CUCK(c, cudaEventRecord(c->ev1, c->stream));
  CUCK(c, cudaEventSynchronize(c->ev1));
  float ms = 0;
  CUCK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->msRender += ms;
  return MOX_OK;
}

void freeGather(mox_ctx* c) {
  for (int w = 0; w < 2; ++w) {
    if (c->dGather[w]) { if (c->gatherPeer[w] == 1) cudaIpcCloseMemHandle(c->dGather[w]); else if (!c->gatherPeer[w]) cudaFree(c->dGather[w]); }
    if (c->hostGather[w]) cudaFreeHost(c->hostGather[w]);
    c->dGather[w] = nullptr; c->hostGather[w] = nullptr; c->gatherPeer[w] = 0;
  }
  c->gatherBytes = 0; c->readPending = -1;
}

// The two gather buffers of this context (allocated on first use, re-allocated when the image size changes).
int ensureGather(mox_ctx* c, bool hostSide) {
  if (!c->dAccu) return fail(c, MOX_ERR_STATE, "no accumulation buffer (mox_set_globals first)");
  const size_t bytes = (size_t)c->accuW * c->accuH * 12;
  if (c->gatherBytes != bytes) {
    cudaStreamSynchronize(c->stream);
    if (c->copyStream) cudaStreamSynchronize(c->copyStream);
    bool peer = c->gatherPeer[0] || c->gatherPeer[1];
    if (peer && c->gatherBytes) return fail(c, MOX_ERR_STATE, "image size changed after mox_gather_import");
    if (!peer) freeGather(c);
    c->gatherBytes = bytes;
  }
  if (!c->copyStream) CUCK(c, cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
  if (!c->evPushed) CUCK(c, cudaEventCreateWithFlags(&c->evPushed, cudaEventDisableTiming));
  for (int w = 0; w < 2; ++w) {
    if (!c->evCopied[w]) CUCK(c, cudaEventCreateWithFlags(&c->evCopied[w], cudaEventDisableTiming));
    if (!c->dGather[w]) { CUCK(c, cudaMalloc(&c->dGather[w], bytes)); CUCK(c, cudaMemsetAsync(c->dGather[w], 0, bytes, c->stream)); }
    if (hostSide && !c->hostGather[w]) CUCK(c, cudaMallocHost(&c->hostGather[w], bytes));
  }
  return MOX_OK;
}

int ensurePinned(mox_ctx* c, size_t bytes) {
  if (c->pinnedBytes >= bytes) return MOX_OK;
  if (c->pinned) cudaFreeHost(c->pinned);
  c->pinned = nullptr; c->pinnedBytes = 0;
  CUCK(c, cudaMallocHost(&c->pinned, bytes));
  c->pinnedBytes = bytes;
  return MOX_OK;
}

}  // namespace

// A multi-GPU handle (mox_create_multi) is a shell: every call is forwarded to its per-device contexts.
#define GROUP_EACH(c, expr)                                                                      \
  do {                                                                                           \
    if ((c)->group) {                                                                            \
      int rc_ = groupEach((c)->group, [&](mox_ctx* k, int) { return (expr); });                  \
      if (rc_) (c)->err = groupError((c)->group);                                                \
      return rc_;                                                                                \
    }                                                                                            \
  } while (0)
#define GROUP_PAR(c, expr)                                                                       \
  do {                                                                                           \
    if ((c)->group) {                                                                            \
      int rc_ = groupParallel((c)->group, [&](mox_ctx* k, int) { return (expr); });              \
      if (rc_) (c)->err = groupError((c)->group);                                                \
      return rc_;                                                                                \
    }                                                                                            \
  } while (0)
#define GROUP_REFUSE(c, what) \
  do { if ((c)->group) return fail((c), MOX_ERR_INVALID, what " is managed by the multi-GPU handle itself"); } while (0)
#define GROUP_FIRST(c, expr)                                                                     \
  do {                                                                                           \
    if ((c)->group) {                                                                            \
      mox_ctx* k = groupChild((c)->group, 0);                                                    \
      int rc_ = (expr);                                                                          \
      if (rc_) (c)->err = mox_last_error(k);                                                     \
      return rc_;                                                                                \
    }                                                                                            \
  } while (0)

extern "C" {

int mox_abi_version(void) { return MOX_ABI_VERSION; }

int mox_create_multi(mox_ctx** out, const int* device_ids, int n_devices) {
  if (!out) return fail(nullptr, MOX_ERR_INVALID, "null out pointer");
  *out = nullptr;
  std::string err;
  mox_group* g = nullptr;
  int rc = groupCreate(&g, device_ids, n_devices, err);
  if (rc) return fail(nullptr, rc, err);
  mox_ctx* shell = new mox_ctx();
  shell->group = g;
  *out = shell;
  return MOX_OK;
}

int mox_device_count(const mox_ctx* c) { return c ? (c->group ? groupCount(c->group) : 1) : 0; }

const char* mox_last_error(const mox_ctx* c) { return c ? c->err.c_str() : g_createError.c_str(); }

int mox_create(mox_ctx** out, int device_id) {
  if (!out) return fail(nullptr, MOX_ERR_INVALID, "null out pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, MOX_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                           " (this library has no CPU fallback)");
  if (device_id < 0 || device_id >= n) return fail(nullptr, MOX_ERR_INVALID, "device_id out of range");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) return fail(nullptr, MOX_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, MOX_ERR_CUDA, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                           "; libmox.so contains sm_100a code only");
  mox_ctx* c = new mox_ctx();
  c->device = device_id;
  if ((e = cudaSetDevice(device_id)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming)) != cudaSuccess) {
    std::string msg = cudaGetErrorString(e);
    delete c;
    return fail(nullptr, MOX_ERR_CUDA, msg);
  }
  if (const char* env = getenv("MOX_SORT_RAYS")) c->sortRays = atoi(env) != 0;
  if (const char* env = getenv("MOX_SORT_PASSES")) c->sortPasses = atoi(env) == 1 ? 1 : 3;
  if (const char* env = getenv("MOX_MAX_BATCH_PATHS")) { long long v = atoll(env); if (v > 0) c->maxBatchPaths = (size_t)v; }
  if (const char* env = getenv("MOX_DISNEY_SPLIT")) c->disneySplit = atoi(env) != 0;
  if (const char* env = getenv("MOX_OVERLAP_SHADOW")) c->overlapShadow = atoi(env) != 0;
  if (const char* env = getenv("MOX_BRDF_IEEE")) c->brdfFast = atoi(env) == 0;
  if (const char* env = getenv("MOX_SLICES")) c->nSlices = std::min(std::max(atoi(env), 1), (int)mox_ctx::kMaxSlices);
  memset(&c->rp, 0, sizeof c->rp);
  c->rp.maxDepth = 256; c->rp.eps = 0.001f; c->rp.minIntensity = 0.001f;
  c->rp.bad = make_float3(1.f, 1.f, 1.f);
  *out = c;
  return MOX_OK;
}

void mox_destroy(mox_ctx* c) {
  if (!c) return;
  if (c->group) { groupDestroy(c->group); delete c; return; }
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (DevBuf* b : {&c->dPrims, &c->dTris, &c->dVerts, &c->dNormals, &c->dUvs, &c->dAnalytic, &c->dMats, &c->dLights, &c->dLightN, &c->dShadeRec, &c->dQueryRays, &c->dQueryOut, &c->dQueryC, &c->dQueryO,
                   &c->dQueryD, &c->dQueryCounters}) b->release();
  for (auto& b : c->otherOwned) b.release();
  freeTextures(c);
  c->dTexObjs.release();
  c->buildArena.release();
  cudaFree(c->dNodes); cudaFree(c->dPacked); cudaFree(c->dNodes8); cudaFree(c->dPacked8); cudaFree(c->dAccu); cudaFree(c->dOwned);
  for (int k = 0; k < mox_ctx::kMaxSlices; ++k) {
    mox_ctx::Slice& sl = c->slices[k];
    if (sl.stream && k) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
    if (sl.shadowStream) { cudaStreamSynchronize(sl.shadowStream); cudaStreamDestroy(sl.shadowStream); }
    if (sl.evShaded) cudaEventDestroy(sl.evShaded);
    if (sl.evApplied) cudaEventDestroy(sl.evApplied);
    freePaths(sl.pb);
    sl.timer.release();
    if (sl.evReady) cudaEventDestroy(sl.evReady);
    if (sl.hostCnt) cudaFreeHost(sl.hostCnt);
  }
  if (c->pinned) cudaFreeHost(c->pinned);
  freeGather(c);
  if (c->copyStream) cudaStreamDestroy(c->copyStream);
  if (c->evPushed) cudaEventDestroy(c->evPushed);
  for (auto e : c->evCopied) if (e) cudaEventDestroy(e);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->evFork) cudaEventDestroy(c->evFork);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int mox_set_globals(mox_ctx* c, uint32_t width, uint32_t height, uint32_t rayMaxDepth, float rayEpsilonT, float rayMinIntensity,
                    const float absorbColor[3], const float badColor[3], const float bgColor[3]) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_EACH(c, mox_set_globals(k, width, height, rayMaxDepth, rayEpsilonT, rayMinIntensity, absorbColor, badColor, bgColor));
  if (!width || !height || !absorbColor || !badColor || !bgColor) return fail(c, MOX_ERR_INVALID, "bad globals");
  if ((uint64_t)width * height > 0x7fffffffull / 3) return fail(c, MOX_ERR_INVALID, "image too large");
  int rc = bind(c);
  if (rc) return rc;
  if (width != c->accuW || height != c->accuH || !c->dAccu) {
    cudaStreamSynchronize(c->stream);
    cudaFree(c->dAccu);
    c->dAccu = nullptr;
    CUCK(c, cudaMalloc(&c->dAccu, (size_t)width * height * 12));
    CUCK(c, cudaMemset(c->dAccu, 0, (size_t)width * height * 12));
    c->accuW = width; c->accuH = height;
    c->launches = 0;
    c->ownedDirty = true;
  }
  c->rp.W = width; c->rp.H = height; c->rp.maxDepth = rayMaxDepth; c->rp.eps = rayEpsilonT; c->rp.minIntensity = rayMinIntensity;
  c->rp.absorb = make_float3(absorbColor[0], absorbColor[1], absorbColor[2]);
  c->rp.bad = make_float3(badColor[0], badColor[1], badColor[2]);
  c->rp.bg = make_float3(bgColor[0], bgColor[1], bgColor[2]);
  c->haveGlobals = true;
  return MOX_OK;
}

int mox_set_camera(mox_ctx* c, const CamParams* cam) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_EACH(c, mox_set_camera(k, cam));
  if (!cam) return fail(c, MOX_ERR_INVALID, "null camera");
  c->rp.cam = *cam;
  c->haveCamera = true;
  return MOX_OK;
}

int mox_set_rng_mode(mox_ctx* c, int mode) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_EACH(c, mox_set_rng_mode(k, mode));
  if (mode != MOX_RNG_REF && mode != MOX_RNG_PHILOX) return fail(c, MOX_ERR_INVALID, "bad rng mode");
  c->rp.rngMode = mode;
  return MOX_OK;
}

int mox_set_partition(mox_ctx* c, uint32_t rank, uint32_t world, uint32_t tile) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_REFUSE(c, "the tile partition");
  if (!world || rank >= world || !tile) return fail(c, MOX_ERR_INVALID, "bad partition");
  c->rank = rank; c->world = world; c->tile = tile;
  c->ownedDirty = true;
  return MOX_OK;
}

int mox_add_texture_rgba32f(mox_ctx* c, const float* texels, int w, int h, int* out_id) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_EACH(c, mox_add_texture_rgba32f(k, texels, w, h, out_id));
  if (!texels || w <= 0 || h <= 0 || w > 32768 || h > 32768) return fail(c, MOX_ERR_INVALID, "bad texture");
  mox_ctx::HostTexture t;
  t.w = w; t.h = h;
  t.texels.assign(texels, texels + (size_t)w * h * 4);
  c->textures.push_back(std::move(t));
  c->texturesDirty = true;
  if (out_id) *out_id = (int)c->textures.size();
  return MOX_OK;
}

int mox_add_sphere(mox_ctx* c, const SphereParams* s, int kind, const void* params, uint32_t* out_id) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_EACH(c, mox_add_sphere(k, s, kind, params, out_id));
  if (!s || !params) return fail(c, MOX_ERR_INVALID, "null argument");
  int m = addMaterial(c, kind, params);
  if (m < 0) return fail(c, MOX_ERR_INVALID, "bad material kind");
  Analytic a;
  memset(&a, 0, sizeof a);
  a.a = make_float4(s->center.x, s->center.y, s->center.z, s->radius);
  c->analytic.push_back(a);
  c->prims.push_back({PT_SPHERE | ((uint32_t)m << 2), (uint32_t)c->analytic.size() - 1});
  c->nSpheres++;
  if (out_id) *out_id = (uint32_t)c->prims.size() - 1;
  c->built = false;
  return MOX_OK;
}

int mox_add_quad(mox_ctx* c, const QuadParams* q, int kind, const void* params, uint32_t* out_id) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_EACH(c, mox_add_quad(k, q, kind, params, out_id));
  if (!q || !params) return fail(c, MOX_ERR_INVALID, "null argument");
  int m = addMaterial(c, kind, params);
  if (m < 0) return fail(c, MOX_ERR_INVALID, "bad material kind");
  Analytic a;
  a.a = make_float4(q->plane.x, q->plane.y, q->plane.z, q->plane.w);
  a.b = make_float4(q->v1.x, q->v1.y, q->v1.z, 0.f);
  a.c = make_float4(q->v2.x, q->v2.y, q->v2.z, 0.f);
  a.d = make_float4(q->anchor.x, q->anchor.y, q->anchor.z, 0.f);
  c->analytic.push_back(a);
  c->prims.push_back({PT_QUAD | ((uint32_t)m << 2), (uint32_t)c->analytic.size() - 1});
  c->nQuads++;
  if (out_id) *out_id = (uint32_t)c->prims.size() - 1;
  c->built = false;
  return MOX_OK;
}

int mox_add_mesh(mox_ctx* c, const float* v, size_t nv, const float* n, size_t nn, const float* uv, size_t nt, const int32_t* vIdx,
                 const int32_t* nIdx, const int32_t* tIdx, size_t nFaces, int kind, const void* params, uint32_t* out_first) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_EACH(c, mox_add_mesh(k, v, nv, n, nn, uv, nt, vIdx, nIdx, tIdx, nFaces, kind, params, out_first));
  if (!params || (nFaces && (!v || !vIdx))) return fail(c, MOX_ERR_INVALID, "null argument");
  if (c->prims.size() + nFaces >= (1u << MOX_HIT_ID_BITS)) return fail(c, MOX_ERR_INVALID, "too many primitives");
  int m = addMaterial(c, kind, params);
  if (m < 0) return fail(c, MOX_ERR_INVALID, "bad material kind");
  bool hasN = nn > 0 && n && nIdx, hasT = nt > 0 && uv && tIdx;
  for (size_t f = 0; f < nFaces * 3; ++f) {
    if (vIdx[f] < 0 || (size_t)vIdx[f] >= nv) { c->mats.pop_back(); return fail(c, MOX_ERR_INVALID, "vertex index out of range"); }
    if (hasN && (nIdx[f] < 0 || (size_t)nIdx[f] >= nn)) hasN = false;
    if (hasT && (tIdx[f] < 0 || (size_t)tIdx[f] >= nt)) hasT = false;
  }
  int vb = (int)(c->verts.size() / 3), nb = (int)(c->normals.size() / 3), tb = (int)(c->uvs.size() / 2);
  c->verts.insert(c->verts.end(), v, v + nv * 3);
  if (hasN) c->normals.insert(c->normals.end(), n, n + nn * 3);
  if (hasT) c->uvs.insert(c->uvs.end(), uv, uv + nt * 2);
  if (out_first) *out_first = (uint32_t)c->prims.size();
  c->tris.reserve(c->tris.size() + nFaces);
  c->prims.reserve(c->prims.size() + nFaces);
  for (size_t f = 0; f < nFaces; ++f) {
    TriIdx t;
    for (int k = 0; k < 3; ++k) {
      t.v[k] = vb + vIdx[3 * f + k];
      t.n[k] = hasN ? nb + nIdx[3 * f + k] : -1;
      t.t[k] = hasT ? tb + tIdx[3 * f + k] : -1;
    }
    c->tris.push_back(t);
    c->prims.push_back({PT_TRI | ((uint32_t)m << 2), (uint32_t)c->tris.size() - 1});
  }
  c->built = false;
  return MOX_OK;
}

int mox_set_lights(mox_ctx* c, const LightParams* l, size_t n) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_EACH(c, mox_set_lights(k, l, n));
  if (n && !l) return fail(c, MOX_ERR_INVALID, "null lights");
  c->lights.assign(l, l + n);
  c->lightsDirty = true;
  return MOX_OK;
}

int mox_clear_scene(mox_ctx* c) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_EACH(c, mox_clear_scene(k));
  c->prims.clear(); c->tris.clear(); c->verts.clear(); c->normals.clear(); c->uvs.clear(); c->analytic.clear();
  c->mats.clear(); c->lights.clear();
  c->textures.clear(); c->texturesDirty = true;
  c->nSpheres = c->nQuads = 0;
  c->built = false; c->lightsDirty = true;
  return MOX_OK;
}

int mox_build_accel(mox_ctx* c, uint32_t flags, float* out_ms) {
  if (!c) return MOX_ERR_INVALID;
  if (c->group) {
    std::vector<float> ms(groupCount(c->group), 0.f);
    int rc_ = groupParallel(c->group, [&](mox_ctx* k, int i) { return mox_build_accel(k, flags, &ms[i]); });
    if (rc_) { c->err = groupError(c->group); return rc_; }
    if (out_ms) *out_ms = *std::max_element(ms.begin(), ms.end());
    return MOX_OK;
  }
  int rc = bind(c);
  if (rc) return rc;
  // A failed rebuild must not leave a context that still claims to be built.
  c->built = false;
  c->wideBuilt = false;
  cudaStreamSynchronize(c->stream);
  if ((rc = upload(c, c->dPrims, c->prims))) return rc;
  if ((rc = upload(c, c->dTris, c->tris))) return rc;
  if ((rc = upload(c, c->dVerts, c->verts))) return rc;
  if ((rc = upload(c, c->dNormals, c->normals))) return rc;
  if ((rc = upload(c, c->dUvs, c->uvs))) return rc;
  if ((rc = upload(c, c->dAnalytic, c->analytic))) return rc;
  if ((rc = upload(c, c->dMats, c->mats))) return rc;
  c->shadeRecBuilt = false;
  {
    bool want = !c->tris.empty();
    if (const char* env = getenv("MOX_SHADE_RECORDS")) want = want && atoi(env) != 0;
    if (want) {
      if ((rc = ensure(c, c->dShadeRec, c->tris.size() * MOX_SHADE_REC_F4 * sizeof(float4)))) return rc;
      launchBuildShadeRecords((const TriIdx*)c->dTris.p, (const float*)c->dVerts.p, (const float*)c->dNormals.p, (const float*)c->dUvs.p,
                              (uint32_t)c->tris.size(), (float4*)c->dShadeRec.p, c->stream);
      c->shadeRecBuilt = true;
    }
  }
  if ((rc = syncLights(c))) return rc;
  if ((rc = syncTextures(c))) return rc;
  c->hasDisneyNormal = false;
  for (auto& m : c->mats) {
    if (m.kind == MOX_MAT_DISNEY && (m.dis.albedoID < 0 || m.dis.albedoID > (int)c->textures.size()))
      return fail(c, MOX_ERR_INVALID, "DisneyParams.albedoID refers to a texture that was not added");
    if (m.kind == MOX_MAT_DISNEY && m.dis.brdfType != GLASS) c->hasDisneyNormal = true;
  }
  BuildInput in;
  in.arena = &c->buildArena;
  in.nPrims = (int)c->prims.size();
  in.prims = (const PrimDesc*)c->dPrims.p;
  in.tris = (const TriIdx*)c->dTris.p;
  in.verts = (const float*)c->dVerts.p;
  in.analytic = (const Analytic*)c->dAnalytic.p;
  in.mats = (const GpuMaterial*)c->dMats.p;
  in.evStart = c->ev0; in.evStop = c->ev1;
  in.usePloc = (flags & MOX_ACCEL_LBVH) == 0;
  if (const char* env = getenv("MOX_PLOC_RADIUS")) in.plocRadius = atoi(env);
  if (const char* env = getenv("MOX_FORCE_LBVH")) { if (atoi(env)) in.usePloc = false; }
  c->useWide = (flags & MOX_ACCEL_BINARY) == 0;
  if (const char* env = getenv("MOX_FORCE_BINARY")) { if (atoi(env)) c->useWide = false; }
  in.useWide = c->useWide;
  if (const char* env = getenv("MOX_WATERTIGHT")) { if (atoi(env)) flags |= MOX_ACCEL_WATERTIGHT; }
  in.watertight = (flags & MOX_ACCEL_WATERTIGHT) != 0;
  BuildOutput out;
  out.nodes = c->dNodes; out.packed = c->dPacked; out.nodesCap = c->nodesCap; out.packedCap = c->packedCap;
  out.nodes8 = c->dNodes8; out.packed8 = c->dPacked8; out.nodes8Cap = c->nodes8Cap; out.packed8Cap = c->packed8Cap;
  std::string err;
  bool okBuild = buildBvh(in, out, c->stream, err);
  c->dNodes = out.nodes; c->dPacked = out.packed; c->nodesCap = out.nodesCap; c->packedCap = out.packedCap;
  c->dNodes8 = out.nodes8; c->dPacked8 = out.packed8; c->nodes8Cap = out.nodes8Cap; c->packed8Cap = out.packed8Cap;
  c->nNodes8 = out.nNodes8; c->wideBuilt = okBuild && out.nNodes8 > 0;
  if (!okBuild) return fail(c, MOX_ERR_CUDA, "build_accel: " + err);
  CUCK(c, cudaEventSynchronize(c->ev1));
  float ms = 0;
  CUCK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->dNodes = out.nodes; c->dPacked = out.packed; c->nNodes = out.nNodes; c->nValid = out.nValid;
  for (int k = 0; k < 3; ++k) { c->sceneLo[k] = out.sceneLo[k]; c->sceneHi[k] = out.sceneHi[k]; }
  c->msBuild = ms;
  c->accelFlags = flags;
  c->built = true;
  if (out_ms) *out_ms = ms;
  return MOX_OK;
}

int mox_launch(mox_ctx* c, int32_t randSeed) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_PAR(c, mox_launch(k, randSeed));
  return renderSeeds(c, std::vector<int32_t>{randSeed});
}

int mox_render(mox_ctx* c, uint32_t spp, uint32_t seed) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_PAR(c, mox_render(k, spp, seed));
  std::vector<int32_t> seeds(spp);
  for (uint32_t i = 0; i < spp; ++i) seeds[i] = (int32_t)tea16((uint32_t)(c->launches + i), seed);
  return renderSeeds(c, seeds);
}

int mox_read_accum(mox_ctx* c, float* dst) {
  if (!c) return MOX_ERR_INVALID;
  if (c->group) {
    const float* p = nullptr;
    int rc_ = mox_read_accum_begin(c);
    if (!rc_) rc_ = mox_read_accum_end(c, &p);
    if (rc_) return rc_;
    if (!dst) return fail(c, MOX_ERR_INVALID, "null destination");
    mox_stats st;
    uint64_t w = 0, h = 0;
    (void)st;
    mox_ctx* k0 = groupChild(c->group, 0);
    w = k0->accuW; h = k0->accuH;
    memcpy(dst, p, (size_t)w * h * 12);
    return MOX_OK;
  }
  if (!dst || !c->dAccu) return fail(c, MOX_ERR_INVALID, "no accumulation buffer");
  int rc = bind(c);
  if (rc) return rc;
  size_t bytes = (size_t)c->accuW * c->accuH * 12;
  if ((rc = ensurePinned(c, bytes))) return rc;
  CUCK(c, cudaMemcpyAsync(c->pinned, c->dAccu, bytes, cudaMemcpyDeviceToHost, c->stream));
  CUCK(c, cudaStreamSynchronize(c->stream));
  memcpy(dst, c->pinned, bytes);
  return MOX_OK;
}

int mox_map_accum(mox_ctx* c, const float** out) {
  if (!c) return MOX_ERR_INVALID;
  if (c->group) {
    int rc_ = mox_read_accum_begin(c);
    return rc_ ? rc_ : mox_read_accum_end(c, out);
  }
  if (!out || !c->dAccu) return fail(c, MOX_ERR_INVALID, "no accumulation buffer");
  int rc = bind(c);
  if (rc) return rc;
  size_t bytes = (size_t)c->accuW * c->accuH * 12;
  if ((rc = ensurePinned(c, bytes))) return rc;
  CUCK(c, cudaMemcpyAsync(c->pinned, c->dAccu, bytes, cudaMemcpyDeviceToHost, c->stream));
  CUCK(c, cudaStreamSynchronize(c->stream));
  *out = c->pinned;
  return MOX_OK;
}

int mox_unmap_accum(mox_ctx* c) { return c ? MOX_OK : MOX_ERR_INVALID; }

// ---- asynchronous read-back and the peer-memory tile gather (see include/mox.h)
float* ctxGatherBuffer(mox_ctx* c, int which) {
  if (bind(c) || ensureGather(c, true)) return nullptr;
  if (c->gatherPeer[which]) { c->err = "gather buffer is imported"; return nullptr; }
  cudaStreamSynchronize(c->stream);   // the zero-fill of a fresh buffer
  return c->dGather[which];
}

int ctxBorrowGatherTarget(mox_ctx* c, int which, float* ptr) {
  if (c->dGather[which] == ptr) return MOX_OK;
  if (bind(c)) return MOX_ERR_CUDA;
  if (c->dGather[which]) {
    cudaStreamSynchronize(c->stream);
    if (c->gatherPeer[which] == 1) cudaIpcCloseMemHandle(c->dGather[which]);
    else if (!c->gatherPeer[which]) cudaFree(c->dGather[which]);
  }
  c->dGather[which] = ptr;
  c->gatherPeer[which] = 2;   // borrowed from another context of this process: never freed here
  c->gatherBytes = (size_t)c->accuW * c->accuH * 12;
  return MOX_OK;
}

int mox_gather_export(mox_ctx* c, int which, void* handle_out) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_REFUSE(c, "the tile gather");
  if (!handle_out || which < 0 || which > 1) return fail(c, MOX_ERR_INVALID, "bad gather_export arguments");
  int rc = bind(c);
  if (rc) return rc;
  if ((rc = ensureGather(c, true))) return rc;
  if (c->gatherPeer[which]) return fail(c, MOX_ERR_STATE, "this rank imported its gather buffer");
  CUCK(c, cudaStreamSynchronize(c->stream));
  static_assert(sizeof(cudaIpcMemHandle_t) == MOX_IPC_HANDLE_BYTES, "IPC handle size");
  CUCK(c, cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle_out, c->dGather[which]));
  return MOX_OK;
}

int mox_gather_import(mox_ctx* c, int which, const void* handle) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_REFUSE(c, "the tile gather");
  if (!handle || which < 0 || which > 1) return fail(c, MOX_ERR_INVALID, "bad gather_import arguments");
  if (!c->dAccu) return fail(c, MOX_ERR_STATE, "mox_gather_import before mox_set_globals");
  int rc = bind(c);
  if (rc) return rc;
  if (c->dGather[which]) {
    cudaStreamSynchronize(c->stream);
    if (c->gatherPeer[which] == 1) cudaIpcCloseMemHandle(c->dGather[which]);
    else if (!c->gatherPeer[which]) cudaFree(c->dGather[which]);
    c->dGather[which] = nullptr;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  void* p = nullptr;
  CUCK(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  c->dGather[which] = (float*)p;
  c->gatherPeer[which] = 1;
  c->gatherBytes = (size_t)c->accuW * c->accuH * 12;
  return MOX_OK;
}

int mox_gather_push(mox_ctx* c, int which) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_REFUSE(c, "the tile gather");
  if (which < 0 || which > 1 || !c->dAccu) return fail(c, MOX_ERR_INVALID, "bad gather_push");
  int rc = bind(c);
  if (rc) return rc;
  if ((rc = refreshOwned(c))) return rc;
  if (!c->dGather[which] && (rc = ensureGather(c, false))) return rc;   // the root (or a single rank) pushes into its own buffer
  launchPushOwned(c->dAccu, c->dOwned, c->nOwned, c->dGather[which], c->stream);
  CUCK(c, cudaStreamSynchronize(c->stream));   // the pixels have landed before a barrier tells the root to read them
  return MOX_OK;
}

int mox_read_gathered_begin(mox_ctx* c, int which) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_REFUSE(c, "the tile gather");
  if (which < 0 || which > 1) return fail(c, MOX_ERR_INVALID, "bad buffer index");
  int rc = bind(c);
  if (rc) return rc;
  if ((rc = ensureGather(c, true))) return rc;
  if (c->gatherPeer[which]) return fail(c, MOX_ERR_STATE, "only the rank that owns the gather buffer reads it");
  // pushes of this rank are ordered by its stream, pushes of the others by the caller's barrier
  CUCK(c, cudaEventRecord(c->evPushed, c->stream));
  CUCK(c, cudaStreamWaitEvent(c->copyStream, c->evPushed, 0));
  CUCK(c, cudaMemcpyAsync(c->hostGather[which], c->dGather[which], c->gatherBytes, cudaMemcpyDeviceToHost, c->copyStream));
  CUCK(c, cudaEventRecord(c->evCopied[which], c->copyStream));
  return MOX_OK;
}

int mox_read_gathered_end(mox_ctx* c, int which, const float** out) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_REFUSE(c, "the tile gather");
  if (which < 0 || which > 1 || !out || !c->hostGather[which]) return fail(c, MOX_ERR_INVALID, "bad read_gathered_end");
  CUCK(c, cudaEventSynchronize(c->evCopied[which]));
  *out = c->hostGather[which];
  return MOX_OK;
}

int mox_read_accum_begin(mox_ctx* c) {
  if (!c) return MOX_ERR_INVALID;
  if (c->group) {
    int rc_ = groupReadBegin(c->group);
    if (rc_) c->err = groupError(c->group);
    return rc_;
  }
  int rc = bind(c);
  if (rc) return rc;
  if ((rc = ensureGather(c, true))) return rc;
  const int which = c->readCur;
  if (c->gatherPeer[which]) return fail(c, MOX_ERR_STATE, "this rank's gather buffers are imported: use mox_gather_push");
  // snapshot on the render stream (the next launch may start right away), copy on the copy stream; the snapshot
  // must not overwrite an image the copy stream is still sending
  CUCK(c, cudaStreamWaitEvent(c->stream, c->evCopied[which], 0));
  CUCK(c, cudaMemcpyAsync(c->dGather[which], c->dAccu, c->gatherBytes, cudaMemcpyDeviceToDevice, c->stream));
  if ((rc = mox_read_gathered_begin(c, which))) return rc;
  c->readPending = which;
  c->readCur ^= 1;
  return MOX_OK;
}

int mox_read_accum_end(mox_ctx* c, const float** out) {
  if (!c) return MOX_ERR_INVALID;
  if (c->group) {
    int rc_ = groupReadEnd(c->group, out);
    if (rc_) c->err = groupError(c->group);
    return rc_;
  }
  if (c->readPending < 0) return fail(c, MOX_ERR_STATE, "mox_read_accum_end without mox_read_accum_begin");
  int rc = mox_read_gathered_end(c, c->readPending, out);
  c->readPending = -1;
  return rc;
}

int mox_clear_accum(mox_ctx* c) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_EACH(c, mox_clear_accum(k));
  int rc = bind(c);
  if (rc) return rc;
  if (c->dAccu) CUCK(c, cudaMemsetAsync(c->dAccu, 0, (size_t)c->accuW * c->accuH * 12, c->stream));
  CUCK(c, cudaStreamSynchronize(c->stream));
  c->launches = 0; c->raysPrimary = c->raysBounce = c->raysShadow = c->nonfinite = c->nodeVisits = c->primTests = 0;
  c->nodeVisitsShadow = c->primTestsShadow = c->raysShadowTraced = c->shadowBlocked = c->shadowTinted = 0;
  c->msRender = 0;
  for (double& m : c->msStage) m = 0;
  for (int d = 0; d < MOX_STATS_DEPTHS; ++d) { c->raysDepth[d] = c->shadowDepth[d] = 0; c->msDepth[0][d] = c->msDepth[1][d] = 0; }
  c->extendLaunches = c->kernelLaunches = 0;
  return MOX_OK;
}

int mox_set_accum(mox_ctx* c, const float* src, uint64_t launches) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_EACH(c, mox_set_accum(k, src, launches));   // every device keeps the pixels it owns
  if (!src || !c->dAccu) return fail(c, MOX_ERR_INVALID, "no accumulation buffer");
  int rc = bind(c);
  if (rc) return rc;
  CUCK(c, cudaMemcpyAsync(c->dAccu, src, (size_t)c->accuW * c->accuH * 12, cudaMemcpyHostToDevice, c->stream));
  CUCK(c, cudaStreamSynchronize(c->stream));
  c->launches = launches;
  return MOX_OK;
}

int mox_update_sphere(mox_ctx* c, uint32_t prim_id, const SphereParams* s) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_EACH(c, mox_update_sphere(k, prim_id, s));
  if (!s || prim_id >= c->prims.size() || (c->prims[prim_id].typeMat & 3u) != PT_SPHERE) return fail(c, MOX_ERR_INVALID, "not a sphere primitive");
  c->analytic[c->prims[prim_id].geom].a = make_float4(s->center.x, s->center.y, s->center.z, s->radius);
  c->built = false;
  return MOX_OK;
}

int mox_owned_pixels(mox_ctx* c, uint32_t rank, uint64_t* out_n) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_REFUSE(c, "the tile partition");
  if (!out_n || rank >= c->world || !c->haveGlobals) return fail(c, MOX_ERR_INVALID, "bad owned_pixels query");
  // closed form: no need to enumerate the pixels
  uint32_t W = c->rp.W, H = c->rp.H, T = c->tile, tx = (W + T - 1) / T, ty = (H + T - 1) / T;
  uint64_t n = 0;
  for (uint32_t j = 0; j < ty; ++j)
    for (uint32_t i = 0; i < tx; ++i)
      if ((i + j) % c->world == rank) n += (uint64_t)(std::min(W, (i + 1) * T) - i * T) * (std::min(H, (j + 1) * T) - j * T);
  *out_n = n;
  return MOX_OK;
}

int mox_pack_owned(mox_ctx* c, void* dev_dst) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_REFUSE(c, "the tile gather");
  if (!dev_dst || !c->dAccu) return fail(c, MOX_ERR_INVALID, "bad pack_owned");
  int rc = bind(c);
  if (rc) return rc;
  if ((rc = refreshOwned(c))) return rc;
  launchPackOwned(c->dAccu, c->dOwned, c->nOwned, (float*)dev_dst, c->stream);
  CUCK(c, cudaStreamSynchronize(c->stream));
  return MOX_OK;
}

int mox_unpack_owned(mox_ctx* c, uint32_t rank, const void* dev_src) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_REFUSE(c, "the tile gather");
  if (!dev_src || !c->dAccu || rank >= c->world) return fail(c, MOX_ERR_INVALID, "bad unpack_owned");
  int rc = bind(c);
  if (rc) return rc;
  if ((rc = refreshOwned(c))) return rc;
  if (c->otherOwned.size() != c->world) { c->otherOwned.resize(c->world); c->ownedCount.assign(c->world, 0); }
  DevBuf& b = c->otherOwned[rank];
  if (!b.p) {  // first gather of this partition: build and upload rank's pixel list once
    std::vector<uint32_t> l;
    ownedList(c->rp.W, c->rp.H, c->tile, c->world, rank, l);
    if ((rc = ensure(c, b, l.size() * 4))) return rc;
    if (!l.empty()) { CUCK(c, cudaMemcpyAsync(b.p, l.data(), l.size() * 4, cudaMemcpyHostToDevice, c->stream)); CUCK(c, cudaStreamSynchronize(c->stream)); }
    c->ownedCount[rank] = l.size();
  }
  launchUnpackOwned(c->dAccu, (const uint32_t*)b.p, (uint32_t)c->ownedCount[rank], (const float*)dev_src, c->stream);
  CUCK(c, cudaStreamSynchronize(c->stream));
  return MOX_OK;
}

int mox_get_stats(mox_ctx* c, mox_stats* s) {
  if (!c) return MOX_ERR_INVALID;
  if (c->group) {
    if (!s) return fail(c, MOX_ERR_INVALID, "null stats");
    int rc_ = groupStats(c->group, s);
    if (rc_) c->err = groupError(c->group);
    return rc_;
  }
  if (!s) return fail(c, MOX_ERR_INVALID, "null stats");
  memset(s, 0, sizeof *s);
  s->rays_primary = c->raysPrimary; s->rays_bounce = c->raysBounce; s->rays_shadow = c->raysShadow;
  s->nonfinite_samples = c->nonfinite; s->launches = c->launches; s->node_visits = c->nodeVisits; s->prim_tests = c->primTests;
  s->ms_render = c->msRender; s->ms_build = c->msBuild;
  s->n_prims = (uint32_t)c->prims.size(); s->n_triangles = (uint32_t)c->tris.size();
  s->n_spheres = c->nSpheres; s->n_quads = c->nQuads;
  const bool wide = c->wideBuilt && c->useWide;
  s->n_nodes = (uint32_t)(wide ? c->nNodes8 : c->nNodes); s->node_bytes = wide ? sizeof(BvhNode8) : sizeof(BvhNode2); s->prim_bytes = 16 * MOX_PACKED_F4;
  s->n_lights = (uint32_t)c->lights.size();
  s->ms_generate = c->msStage[ST_GENERATE]; s->ms_extend = c->msStage[ST_EXTEND]; s->ms_shade = c->msStage[ST_SHADE];
  s->ms_shadow = c->msStage[ST_SHADOW]; s->ms_accumulate = c->msStage[ST_ACCUMULATE];
  s->extend_launches = c->extendLaunches; s->kernel_launches = c->kernelLaunches;
  s->node_visits_shadow = c->nodeVisitsShadow; s->prim_tests_shadow = c->primTestsShadow;
  s->rays_shadow_traced = c->raysShadowTraced;
  s->rays_shadow_blocked = c->shadowBlocked; s->rays_shadow_tinted = c->shadowTinted;
  for (int d = 0; d < MOX_STATS_DEPTHS; ++d) {
    s->rays_depth[d] = c->raysDepth[d]; s->shadow_traced_depth[d] = c->shadowDepth[d];
    s->ms_extend_depth[d] = c->msDepth[0][d]; s->ms_shadow_depth[d] = c->msDepth[1][d];
  }
  return MOX_OK;
}

int mox_get_device_stats(mox_ctx* c, int index, mox_stats* s) {
  if (!c) return MOX_ERR_INVALID;
  if (index < 0 || index >= mox_device_count(c)) return fail(c, MOX_ERR_INVALID, "device index out of range");
  return mox_get_stats(c->group ? groupChild(c->group, index) : c, s);
}

int mox_trace_closest_device(mox_ctx* c, const void* dev_rays, size_t n, void* dev_hits, float* out_ms) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_FIRST(c, mox_trace_closest_device(k, dev_rays, n, dev_hits, out_ms));
  if (n && (!dev_rays || !dev_hits)) return fail(c, MOX_ERR_INVALID, "null argument");
  if (n > 0xfffffff0ull) return fail(c, MOX_ERR_INVALID, "too many rays");
  if (!c->built) return fail(c, MOX_ERR_STATE, "trace before mox_build_accel");
  if (out_ms) *out_ms = 0.f;
  if (!n) return MOX_OK;
  int rc = bind(c);
  if (rc) return rc;
  bool count = (c->accelFlags & MOX_ACCEL_COUNTERS) != 0;
  if ((rc = ensure(c, c->dQueryO, n * 16))) return rc;
  if ((rc = ensure(c, c->dQueryD, n * 16))) return rc;
  if ((rc = ensure(c, c->dQueryCounters, C_WORDS * 4))) return rc;
  uint32_t* counters = (uint32_t*)c->dQueryCounters.p;
  CUCK(c, cudaMemsetAsync(counters, 0, C_WORDS * 4, c->stream));
  launchSplitRays((const float4*)dev_rays, (float4*)c->dQueryO.p, (float4*)c->dQueryD.p, n, c->stream);
  TraceJob job;
  job.rayO = (const float4*)c->dQueryO.p; job.rayD = (const float4*)c->dQueryD.p; job.queue = nullptr; job.count = (uint32_t)n; job.countPtr = nullptr; job.originMod = 0;
  job.cursor = counters + C_CURSOR; job.hits = (float4*)dev_hits; job.hits2 = nullptr; job.shC = nullptr; job.counters = counters;
  CUCK(c, cudaMemsetAsync(job.cursor, 0, 4, c->stream));
  CUCK(c, cudaEventRecord(c->ev0, c->stream));
  launchTraverse(sceneView(c), job, false, count, c->stream);
  CUCK(c, cudaEventRecord(c->ev1, c->stream));
  CUCK(c, cudaEventSynchronize(c->ev1));
  CUCK(c, cudaGetLastError());
  float ms = 0;
  CUCK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  if (out_ms) *out_ms = ms;
  if (count) {
    uint32_t host[C_WORDS];
    CUCK(c, cudaMemcpy(host, counters, sizeof host, cudaMemcpyDeviceToHost));
    c->nodeVisits += ((uint64_t)host[C_NODEVIS_HI] << 32) | host[C_NODEVIS_LO];
    c->primTests += ((uint64_t)host[C_PRIMTEST_HI] << 32) | host[C_PRIMTEST_LO];
  }
  return MOX_OK;
}

int mox_trace_closest(mox_ctx* c, const float* rays, size_t n, void* hits) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_FIRST(c, mox_trace_closest(k, rays, n, hits));
  if (n && (!rays || !hits)) return fail(c, MOX_ERR_INVALID, "null argument");
  if (!c->built) return fail(c, MOX_ERR_STATE, "trace before mox_build_accel");
  if (!n) return MOX_OK;
  int rc = bind(c);
  if (rc) return rc;
  if ((rc = ensure(c, c->dQueryRays, n * 32)) || (rc = ensure(c, c->dQueryOut, n * 16))) return rc;
  void *dRays = c->dQueryRays.p, *dHits = c->dQueryOut.p;
  // stream-ordered: the traversal kernel runs on the (non-blocking) context stream
  CUCK(c, cudaMemcpyAsync(dRays, rays, n * 32, cudaMemcpyHostToDevice, c->stream));
  if ((rc = mox_trace_closest_device(c, dRays, n, dHits, nullptr))) return rc;
  CUCK(c, cudaMemcpy(hits, dHits, n * 16, cudaMemcpyDeviceToHost));
  return MOX_OK;
}

int mox_trace_shadow(mox_ctx* c, const float* rays, size_t n, float* out_rgb) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_FIRST(c, mox_trace_shadow(k, rays, n, out_rgb));
  if (n && (!rays || !out_rgb)) return fail(c, MOX_ERR_INVALID, "null argument");
  if (n > 0xfffffff0ull) return fail(c, MOX_ERR_INVALID, "too many rays");
  if (!c->built) return fail(c, MOX_ERR_STATE, "trace before mox_build_accel");
  if (!n) return MOX_OK;
  int rc = bind(c);
  if (rc) return rc;
  if ((rc = ensure(c, c->dQueryRays, n * 32)) || (rc = ensure(c, c->dQueryOut, n * 16)) || (rc = ensure(c, c->dQueryC, n * 16)) ||
      (rc = ensure(c, c->dQueryO, n * 16)) || (rc = ensure(c, c->dQueryD, n * 16)) || (rc = ensure(c, c->dQueryCounters, C_WORDS * 4)))
    return rc;
  void *dRays = c->dQueryRays.p, *dOut = c->dQueryOut.p, *dC = c->dQueryC.p;
  uint32_t* counters = (uint32_t*)c->dQueryCounters.p;
  cudaMemcpyAsync(dRays, rays, n * 32, cudaMemcpyHostToDevice, c->stream);   // stream-ordered, see mox_trace_closest
  cudaMemsetAsync(counters, 0, C_WORDS * 4, c->stream);
  launchSplitRays((const float4*)dRays, (float4*)c->dQueryO.p, (float4*)c->dQueryD.p, n, c->stream);
  launchFillOnes((float4*)dC, n, c->stream);
  TraceJob job;
  job.rayO = (const float4*)c->dQueryO.p; job.rayD = (const float4*)c->dQueryD.p; job.queue = nullptr; job.count = (uint32_t)n; job.countPtr = nullptr; job.originMod = 0;
  job.cursor = counters + C_CURSOR; job.hits = nullptr; job.hits2 = nullptr; job.shC = (float4*)dC; job.counters = counters;
  launchTraverse(sceneView(c), job, true, false, c->stream);
  launchCopyRgb((const float4*)dC, (float*)dOut, n, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out_rgb, dOut, n * 12, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail(c, MOX_ERR_CUDA, cudaGetErrorString(e));
  return MOX_OK;
}

}  // extern "C"

// ---- debug / test hooks (not part of the drop-in surface; declared in include/mox_debug.h)
extern "C" int mox_debug_radix_sort(mox_ctx* c, uint32_t* keys, uint32_t* vals, size_t n) {
  if (!c) return MOX_ERR_INVALID;
  GROUP_FIRST(c, mox_debug_radix_sort(k, keys, vals, n));
  if (n && (!keys || !vals)) return fail(c, MOX_ERR_INVALID, "null argument");
  if (!n) return MOX_OK;
  int rc = bind(c);
  if (rc) return rc;
  uint32_t *dk = nullptr, *dv = nullptr;
  CUCK(c, cudaMalloc(&dk, n * 4));
  CUCK(c, cudaMalloc(&dv, n * 4));
  CUCK(c, cudaMemcpyAsync(dk, keys, n * 4, cudaMemcpyHostToDevice, c->stream));
  CUCK(c, cudaMemcpyAsync(dv, vals, n * 4, cudaMemcpyHostToDevice, c->stream));
  std::string err;
  bool ok = radixSortPairs(dk, dv, (int)n, c->stream, err);
  if (ok) {
    cudaMemcpy(keys, dk, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(vals, dv, n * 4, cudaMemcpyDeviceToHost);
  }
  cudaFree(dk); cudaFree(dv);
  return ok ? MOX_OK : fail(c, MOX_ERR_CUDA, err);
}
