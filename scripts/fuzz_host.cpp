// Mutation fuzzer for the host-side readers (images, .scene + OBJ, accumulator dumps) under ASan/UBSan.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <random>
#include <fstream>
#include <dirent.h>
#include <sys/stat.h>
#include "mox_host.h"
static std::vector<uint8_t> slurp(const std::string& p) { std::ifstream f(p, std::ios::binary); return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), {}); }
static void spit(const std::string& p, const std::vector<uint8_t>& d) { std::ofstream f(p, std::ios::binary); f.write((const char*)d.data(), d.size()); }
static void mutate(std::vector<uint8_t>& d, std::mt19937& g) {
  if (d.empty()) return;
  int n = 1 + g() % 8;
  for (int i = 0; i < n; ++i) {
    size_t pos = g() % d.size();
    switch (g() % 6) {
      case 0: d[pos] = (uint8_t)g(); break;
      case 1: d[pos] ^= 1u << (g() % 8); break;
      case 2: d.resize(pos); if (d.empty()) d.push_back(0); break;                       // truncate
      case 3: { size_t len = 1 + g() % 16; d.insert(d.begin() + pos, len, (uint8_t)g()); } break;
      case 4: { size_t len = std::min<size_t>(1 + g() % 16, d.size() - pos); d.erase(d.begin() + pos, d.begin() + pos + len); if (d.empty()) d.push_back(0); } break;
      case 5: { const char* tok[] = {"-1", "0", "999999999", "4000000000", "1e999", "nan", "-", "/", "//", "f 1 2 3 4 5 6 7 8 9", "f -1 -2 -3", "f 1/1/1 2/2/2 3/3/3", "\n", "{", "}", "mesh", "light", "material"};
                const char* t = tok[g() % (sizeof tok / sizeof *tok)]; d.insert(d.begin() + pos, t, t + strlen(t)); } break;
    }
  }
}
int main(int argc, char** argv) {
  // argv: mode (image|scene) seed iterations workdir inputs...
  std::string mode = argv[1]; unsigned seed = atoi(argv[2]); int iters = atoi(argv[3]); std::string work = argv[4];
  std::mt19937 g(seed);
  std::vector<std::string> inputs(argv + 5, argv + argc);
  mkdir(work.c_str(), 0755);
  long ok = 0;
  for (int it = 0; it < iters; ++it) {
    if (mode == "image") {
      const std::string& in = inputs[g() % inputs.size()];
      auto d = slurp(in);
      mutate(d, g);
      std::string ext = in.substr(in.rfind('.'));
      std::string p = work + "/m" + ext;
      spit(p, d);
      int w = 0, h = 0; float* tex = nullptr;
      if (moxh_read_image(p.c_str(), &w, &h, &tex) == 0) { volatile float s = 0; for (long i = 0; i < (long)w * h * 4; i += 97) s += tex[i]; moxh_free(tex); }
    } else {
      // scene dir: copy all files, mutate one of them
      std::string sdir = work + "/s";
      mkdir(sdir.c_str(), 0755); mkdir((sdir + "/x").c_str(), 0755);
      size_t pick = g() % inputs.size();
      std::string sceneName;
      for (size_t k = 0; k < inputs.size(); ++k) {
        auto d = slurp(inputs[k]);
        if (k == pick) mutate(d, g);
        std::string base = inputs[k].substr(inputs[k].rfind('/') + 1);
        if (base.size() > 6 && base.substr(base.size() - 6) == ".scene") base = "x.scene";
        spit(sdir + "/x/" + base, d);
      }
      moxh_scene* s = nullptr;
      if (moxh_scene_load((sdir + "/x").c_str(), "x", &s) == 0 && s) {
        ++ok;
        moxh_scene_info info; moxh_scene_get_info(s, &info);
        CamParams cp; moxh_scene_cam_params(s, 64, 64, &cp);
        uint64_t h4[4]; for (uint32_t m = 0; m < 64; ++m) if (moxh_scene_mesh_hash(s, m, h4) != 0) break;
        moxh_scene_free(s);
      }
    }
  }
  printf("done %s seed %u iters %d loaded %ld\n", mode.c_str(), seed, iters, ok);
  return 0;
}
