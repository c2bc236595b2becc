// cli_main.cpp — headless driver: the replacement of the Qt application
// (MinimalOptiX::MinimalOptiX / renderScene / imageDemo, MinimalOptiX.cpp:9-33, 86-111,
// 540-560).  Loads libmox.so (the GPU library; there is no other backend for the product),
// uploads a scene, renders spp samples with the power-of-two snapshot schedule of
// renderScene(autoSave=true) and writes PNGs plus a JSON stats line.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "api_table.h"
#include "mox_host.h"
#include "scene_desc.h"

static void usage() {
  fprintf(stderr,
          "usage: mox_cli --scene NAME [--scene-dir DIR] [--width W --height H] [--spp N] [--max-depth D]\n"
          "               [--seed S] [--rng ref|philox] [--out PREFIX] [--snapshots] [--dump-accum] [--resume FILE.moxa] [--device K | --gpus N]\n"
          "               [--param P] [--lib PATH] [--watertight]\n"
          "  NAME: spheres_lens spheres_pinhole random_spheres interior soup, or a folder under DIR\n"
          "        holding NAME.scene (coffee, cornell, ...).  Defaults are the reference's constants\n"
          "        (1920x1080, 32 spp, depth 256; MinimalOptiX.h:82-89); a .scene file's properties{width,height}\n"
          "        replace 1920x1080.  --gpus N renders on devices 0..N-1 of this process (tile-split, scene\n"
          "        replicated, tiles gathered over NVLink into device 0).  --watertight: MOX_ACCEL_WATERTIGHT (include/mox.h).\n");
}

int main(int argc, char** argv) {
  std::string scene = "spheres_lens", dir = "scenes", out = "out", lib, rng = "ref", resume;
  uint32_t W = 0, H = 0, spp = 32, depth = 256, seed = 0xC0FFEE;
  uint64_t param = 0;
  int device = 0, gpus = 0;
  bool snapshots = false, dumpAccum = false, watertight = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&]() -> const char* { if (i + 1 >= argc) { usage(); exit(2); } return argv[++i]; };
    if (a == "--scene") scene = val();
    else if (a == "--scene-dir") dir = val();
    else if (a == "--width") W = (uint32_t)atoi(val());
    else if (a == "--height") H = (uint32_t)atoi(val());
    else if (a == "--spp") spp = (uint32_t)atoi(val());
    else if (a == "--max-depth") depth = (uint32_t)atoi(val());
    else if (a == "--seed") seed = (uint32_t)strtoul(val(), nullptr, 0);
    else if (a == "--rng") rng = val();
    else if (a == "--out") out = val();
    else if (a == "--snapshots") snapshots = true;
    else if (a == "--dump-accum") dumpAccum = true;
    else if (a == "--watertight") watertight = true;
    else if (a == "--resume") resume = val();
    else if (a == "--device") device = atoi(val());
    else if (a == "--gpus") gpus = atoi(val());
    else if (a == "--param") param = strtoull(val(), nullptr, 0);
    else if (a == "--lib") lib = val();
    else { usage(); return 2; }
  }
  if (lib.empty()) {
    std::string self = argv[0];
    size_t p = self.rfind('/');
    lib = (p == std::string::npos ? std::string(".") : self.substr(0, p)) + "/libmox.so";
  }
  MoxApi api;
  std::string err;
  if (!loadMoxApi(lib.c_str(), "mox_", api, err)) { fprintf(stderr, "cannot load %s: %s\n", lib.c_str(), err.c_str()); return 1; }
  moxh::SceneDesc sc;
  bool ok;
  if (scene == "spheres_lens" || scene == "spheres_pinhole") ok = moxh::buildSpheres(sc, scene == "spheres_pinhole", 0, 0);
  else if (scene == "random_spheres") ok = moxh::buildRandomSpheres(sc, param ? (int)param : 256, 42u);
  else if (scene == "interior") ok = moxh::buildInterior(sc, param ? param : 1000000ull, 0xD1A1A6u);
  else if (scene == "soup") ok = moxh::buildSoup(sc, param ? param : 10000000ull, 10000000ull);
  else ok = moxh::loadSceneFile(sc, dir + "/" + scene, scene, err);
  if (!ok) { fprintf(stderr, "scene: %s\n", err.c_str()); return 1; }
  for (auto& w : sc.warnings) fprintf(stderr, "warning: %s\n", w.c_str());
  if (!W) W = sc.defaultWidth;
  if (!H) H = sc.defaultHeight;

  mox_ctx* ctx = nullptr;
  if (gpus > 0) {
    if (!api.create_multi) { fprintf(stderr, "%s has no mox_create_multi\n", lib.c_str()); return 1; }
    std::vector<int> ids(gpus);
    for (int i = 0; i < gpus; ++i) ids[i] = i;
    if (api.create_multi(&ctx, ids.data(), gpus)) { fprintf(stderr, "mox_create_multi: %s\n", api.last_error(nullptr)); return 1; }
  } else if (api.create(&ctx, device)) { fprintf(stderr, "mox_create: %s\n", api.last_error(nullptr)); return 1; }
  api.set_rng_mode(ctx, rng == "philox" ? MOX_RNG_PHILOX : MOX_RNG_REF);
  if (!moxh::uploadScene(sc, api, ctx, W, H, depth, err)) { fprintf(stderr, "upload: %s\n", err.c_str()); return 1; }
  float buildMs = 0;
  if (api.build_accel(ctx, watertight ? MOX_ACCEL_WATERTIGHT : MOX_ACCEL_DEFAULT, &buildMs)) { fprintf(stderr, "build_accel: %s\n", api.last_error(ctx)); return 1; }

  std::vector<float> accum((size_t)W * H * 3);
  std::vector<uint8_t> rgb((size_t)W * H * 3);
  // a snapshot = updateContent (MinimalOptiX.cpp:556-560): map the accumulation buffer (on several GPUs: gather the
  // tiles into device 0 over NVLink), quantise, write the image
  double readSec = 0, writeSec = 0;
  uint32_t nSnapshots = 0;
  auto save = [&](const std::string& name, uint32_t n) {
    auto a0 = std::chrono::steady_clock::now();
    api.read_accum(ctx, accum.data());
    auto a1 = std::chrono::steady_clock::now();
    moxh_accum_to_rgb8(accum.data(), W, H, (float)n, rgb.data());
    if (moxh_write_image((name + ".png").c_str(), rgb.data(), W, H)) fprintf(stderr, "write: %s\n", moxh_last_error());
    readSec += std::chrono::duration<double>(a1 - a0).count();
    writeSec += std::chrono::duration<double>(std::chrono::steady_clock::now() - a1).count();
    nSnapshots++;
  };
  uint32_t done = 0, checkpoint = 1;
  if (!resume.empty()) {  // continue a previous run: same seed schedule, samples [launches, spp)
    uint64_t launches = 0;
    if (moxh_read_accum(resume.c_str(), accum.data(), W, H, &launches)) { fprintf(stderr, "resume: %s\n", moxh_last_error()); return 1; }
    if (api.set_accum(ctx, accum.data(), launches)) { fprintf(stderr, "set_accum: %s\n", api.last_error(ctx)); return 1; }
    done = (uint32_t)std::min<uint64_t>(launches, spp);
    while (checkpoint <= done) checkpoint *= 2;
  }
  auto t0 = std::chrono::steady_clock::now();
  while (done < spp) {  // renderScene: snapshots at 1, 2, 4, ... (MinimalOptiX.cpp:543-553)
    uint32_t n = snapshots ? std::min(spp, checkpoint) - done : spp - done;
    if (api.render(ctx, n, seed)) { fprintf(stderr, "render: %s\n", api.last_error(ctx)); return 1; }
    done += n;
    if (snapshots && done == checkpoint) { save(out + "_" + std::to_string(done), done); checkpoint *= 2; }
  }
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  save(out, spp);
  if (dumpAccum) moxh_write_accum((out + ".moxa").c_str(), accum.data(), W, H, spp);
  mox_stats st;
  api.get_stats(ctx, &st);
  double rays = (double)(st.rays_primary + st.rays_bounce);
  printf("{\"scene\": \"%s\", \"width\": %u, \"height\": %u, \"spp\": %u, \"max_depth\": %u, \"seed\": %u, \"rng\": \"%s\", "
         "\"triangles\": %u, \"prims\": %u, \"bvh_build_ms\": %.3f, \"render_ms\": %.3f, \"wall_s\": %.3f, "
         "\"rays_primary\": %llu, \"rays_bounce\": %llu, \"rays_shadow\": %llu, \"mrays_per_s\": %.2f, \"mshadow_per_s\": %.2f, "
         "\"spp_per_s\": %.3f, \"nonfinite\": %llu, \"gpus\": %d, \"ms_extend\": %.2f, \"ms_shade\": %.2f, \"ms_shadow\": %.2f, "
         "\"kernel_launches\": %llu, \"snapshots\": %u, \"gather_and_readback_s\": %.4f, \"quantise_and_png_s\": %.3f, "
         "\"render_ms_per_gpu\": [",
         scene.c_str(), W, H, spp, depth, seed, rng.c_str(), st.n_triangles, st.n_prims, buildMs, st.ms_render, sec,
         (unsigned long long)st.rays_primary, (unsigned long long)st.rays_bounce, (unsigned long long)st.rays_shadow,
         rays / (st.ms_render * 1e3), (double)st.rays_shadow / (st.ms_render * 1e3), spp / (st.ms_render * 1e-3),
         (unsigned long long)st.nonfinite_samples, gpus > 0 ? gpus : 1, st.ms_extend, st.ms_shade, st.ms_shadow,
         (unsigned long long)st.kernel_launches, nSnapshots, readSec, writeSec);
  const int nDev = api.device_count ? api.device_count(ctx) : 1;
  for (int i = 0; i < nDev; ++i) {
    mox_stats ds = st;
    if (api.get_device_stats) api.get_device_stats(ctx, i, &ds);
    printf("%s%.2f", i ? ", " : "", ds.ms_render);
  }
  printf("]}\n");
  api.destroy(ctx);
  return 0;
}

