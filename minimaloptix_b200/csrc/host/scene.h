// scene.h — the `.scene` description loader (reference: MinimalOptiX/scene.h:18-27,
// scene.cpp:5-124; grammar in SURVEY.md Appendix B.4).  Same class name and public members
// as the reference so host code that consumes a Scene keeps working; written from scratch
// around a small key/value line matcher instead of the reference's sscanf ladder.
//
// Accepted input is the same; behaviour differs only where the reference has a bug
// (SURVEY.md Appendix C): a missing file or an unknown material name sets ok=false and
// `error` instead of crashing / desynchronising the vectors (Q12, Q13), and LightParams are
// zero-initialised (Q11).
#pragma once
#include <string>
#include <vector>
#include "mox_structs.h"

class Scene {
 public:
  explicit Scene(const char* fileName);
  std::vector<std::string> meshNames;   // one entry per mesh{} block
  std::vector<DisneyParams> materials;  // parallel to meshNames
  std::vector<std::string> textures;    // parallel to meshNames ("" = none)
  std::vector<LightParams> lights;
  int width = 0;
  int height = 0;
  // additions
  bool ok = true;
  std::string error;
};
