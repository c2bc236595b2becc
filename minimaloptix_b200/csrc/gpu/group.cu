// group.cu — mox_create_multi: one handle, several GPUs of this process.
//
// The reference renders on one device through one optix::Context (MinimalOptiX.cpp:131).  Here the path
// shards by pixels (Camera.cu:24 seeds by pixel index only): child i of n renders the 32x32 tiles with
// (tx + ty) % n == i of a replicated scene, one persistent host thread per device drives its child
// through the ordinary C ABI, and a read-back lets every device write its own pixels straight into
// device 0's gather buffer with peer stores over NVLink (cudaDeviceEnablePeerAccess) — no staging
// buffer, no collective, no host in the data path — before device 0 copies the frame to the host.
#include "group.h"

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

namespace {

// One host thread per device, alive for the life of the group.
struct Worker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<int()> task;
  bool hasTask = false, done = false, quit = false;
  int rc = 0;
  void loop() {
    std::unique_lock<std::mutex> lk(m);
    for (;;) {
      cv.wait(lk, [&] { return hasTask || quit; });
      if (quit) return;
      std::function<int()> t = std::move(task);
      hasTask = false;
      lk.unlock();
      int r = t();
      lk.lock();
      rc = r; done = true;
      cv.notify_all();
    }
  }
  void post(std::function<int()> t) {
    std::lock_guard<std::mutex> lk(m);
    task = std::move(t); hasTask = true; done = false;
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return done; });
    return rc;
  }
};

}  // namespace

struct mox_group {
  std::vector<mox_ctx*> kids;
  std::vector<int> devices;
  std::vector<Worker*> workers;
  std::string err;
  int readCur = 0, readPending = -1;
  bool targetsSet = false;
};

int groupCreate(mox_group** out, const int* deviceIds, int n, std::string& err) {
  *out = nullptr;
  if (!deviceIds || n < 1 || n > 64) { err = "bad device list"; return MOX_ERR_INVALID; }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j)
      if (deviceIds[i] == deviceIds[j]) { err = "device listed twice"; return MOX_ERR_INVALID; }
  mox_group* g = new mox_group();
  for (int i = 0; i < n; ++i) {
    mox_ctx* k = nullptr;
    int rc = mox_create(&k, deviceIds[i]);
    if (rc) { err = mox_last_error(nullptr); groupDestroy(g); return rc; }
    g->kids.push_back(k);
    g->devices.push_back(deviceIds[i]);
    mox_set_partition(k, (uint32_t)i, (uint32_t)n, 32);
  }
  // every device may store into device 0's memory
  for (int i = 1; i < n; ++i) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, deviceIds[i], deviceIds[0]);
    if (!can) { err = "device " + std::to_string(deviceIds[i]) + " cannot access device " + std::to_string(deviceIds[0]) + " (no peer-to-peer path)"; groupDestroy(g); return MOX_ERR_CUDA; }
    cudaSetDevice(deviceIds[i]);
    cudaError_t e = cudaDeviceEnablePeerAccess(deviceIds[0], 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); groupDestroy(g); return MOX_ERR_CUDA; }
    cudaGetLastError();
  }
  for (int i = 0; i < n; ++i) {
    Worker* w = new Worker();
    w->th = std::thread([w] { w->loop(); });
    g->workers.push_back(w);
  }
  *out = g;
  return MOX_OK;
}

void groupDestroy(mox_group* g) {
  if (!g) return;
  for (Worker* w : g->workers) {
    { std::lock_guard<std::mutex> lk(w->m); w->quit = true; w->cv.notify_all(); }
    w->th.join();
    delete w;
  }
  // children first release what they borrowed from child 0, then child 0 goes last
  for (size_t i = g->kids.size(); i-- > 0;) mox_destroy(g->kids[i]);
  delete g;
}

int groupCount(const mox_group* g) { return (int)g->kids.size(); }
mox_ctx* groupChild(mox_group* g, int i) { return g->kids[i]; }
const std::string& groupError(const mox_group* g) { return g->err; }

int groupEach(mox_group* g, const std::function<int(mox_ctx*, int)>& f) {
  for (size_t i = 0; i < g->kids.size(); ++i) {
    int rc = f(g->kids[i], (int)i);
    if (rc) { g->err = "device " + std::to_string(g->devices[i]) + ": " + mox_last_error(g->kids[i]); return rc; }
  }
  return MOX_OK;
}

int groupParallel(mox_group* g, const std::function<int(mox_ctx*, int)>& f) {
  const size_t n = g->kids.size();
  if (n == 1) return groupEach(g, f);
  for (size_t i = 0; i < n; ++i) {
    mox_ctx* k = g->kids[i];
    int idx = (int)i;
    g->workers[i]->post([&f, k, idx] { return f(k, idx); });
  }
  int first = MOX_OK;
  for (size_t i = 0; i < n; ++i) {
    int rc = g->workers[i]->wait();
    if (rc && !first) { first = rc; g->err = "device " + std::to_string(g->devices[i]) + ": " + mox_last_error(g->kids[i]); }
  }
  return first;
}

// Gather + start of the device->host copy.  All devices push at once; pushes are synchronous in their
// host thread, so when groupParallel returns the frame is complete in device 0's buffer.
int groupReadBegin(mox_group* g) {
  const int which = g->readCur;
  float* target[2] = {ctxGatherBuffer(g->kids[0], 0), ctxGatherBuffer(g->kids[0], 1)};
  if (!target[0] || !target[1]) { g->err = mox_last_error(g->kids[0]); return MOX_ERR_STATE; }
  for (size_t i = 1; i < g->kids.size(); ++i)
    for (int w = 0; w < 2; ++w) {
      int rc = ctxBorrowGatherTarget(g->kids[i], w, target[w]);
      if (rc) { g->err = mox_last_error(g->kids[i]); return rc; }
    }
  int rc = groupParallel(g, [which](mox_ctx* k, int) { return mox_gather_push(k, which); });
  if (rc) return rc;
  if ((rc = mox_read_gathered_begin(g->kids[0], which))) { g->err = mox_last_error(g->kids[0]); return rc; }
  g->readPending = which;
  g->readCur ^= 1;
  return MOX_OK;
}

int groupReadEnd(mox_group* g, const float** out) {
  if (g->readPending < 0) { g->err = "mox_read_accum_end without mox_read_accum_begin"; return MOX_ERR_STATE; }
  int rc = mox_read_gathered_end(g->kids[0], g->readPending, out);
  if (rc) g->err = mox_last_error(g->kids[0]);
  g->readPending = -1;
  return rc;
}

int groupStats(mox_group* g, mox_stats* out) {
  mox_stats acc;
  int rc = mox_get_stats(g->kids[0], &acc);
  if (rc) return rc;
  for (size_t i = 1; i < g->kids.size(); ++i) {
    mox_stats s;
    if ((rc = mox_get_stats(g->kids[i], &s))) return rc;
    acc.rays_primary += s.rays_primary; acc.rays_bounce += s.rays_bounce; acc.rays_shadow += s.rays_shadow;
    acc.nonfinite_samples += s.nonfinite_samples; acc.node_visits += s.node_visits; acc.prim_tests += s.prim_tests;
    acc.node_visits_shadow += s.node_visits_shadow; acc.prim_tests_shadow += s.prim_tests_shadow;
    acc.rays_shadow_blocked += s.rays_shadow_blocked; acc.rays_shadow_tinted += s.rays_shadow_tinted;
    for (int d = 0; d < MOX_STATS_DEPTHS; ++d) {
      acc.rays_depth[d] += s.rays_depth[d]; acc.shadow_traced_depth[d] += s.shadow_traced_depth[d];
      acc.ms_extend_depth[d] = acc.ms_extend_depth[d] > s.ms_extend_depth[d] ? acc.ms_extend_depth[d] : s.ms_extend_depth[d];
      acc.ms_shadow_depth[d] = acc.ms_shadow_depth[d] > s.ms_shadow_depth[d] ? acc.ms_shadow_depth[d] : s.ms_shadow_depth[d];
    }
    acc.rays_shadow_traced += s.rays_shadow_traced; acc.extend_launches += s.extend_launches; acc.kernel_launches += s.kernel_launches;
    // devices run side by side: times are the slowest device's
    acc.ms_render = acc.ms_render > s.ms_render ? acc.ms_render : s.ms_render;
    acc.ms_build = acc.ms_build > s.ms_build ? acc.ms_build : s.ms_build;
    acc.ms_generate = acc.ms_generate > s.ms_generate ? acc.ms_generate : s.ms_generate;
    acc.ms_extend = acc.ms_extend > s.ms_extend ? acc.ms_extend : s.ms_extend;
    acc.ms_shade = acc.ms_shade > s.ms_shade ? acc.ms_shade : s.ms_shade;
    acc.ms_shadow = acc.ms_shadow > s.ms_shadow ? acc.ms_shadow : s.ms_shadow;
    acc.ms_accumulate = acc.ms_accumulate > s.ms_accumulate ? acc.ms_accumulate : s.ms_accumulate;
  }
  *out = acc;
  return MOX_OK;
}
