// Compile-and-link check of the C++ mirror class against the shared library (no GPU needed to build; at run time it
// exercises the loaders and the error paths; detection itself is covered by the gpu tests through the same C ABI).
#include <cstdio>
#include <cstring>
#include "cylindertag/CylinderTag.h"
using namespace ctag_api;
int main(int argc, char** argv) {
  if (argc < 4) return 2;
  std::vector<ModelInfo> models;
  CamInfo cam;
  int created = 0;
  try {
    CylinderTag tag(argv[1]);
    created = 1;
    tag.loadModel(argv[2], models);
    tag.loadCamera(argv[3], cam);
  } catch (const std::string& s) {
    std::printf("exception: %s", s.c_str());
  }
  try {
    CylinderTag bad("/nonexistent.marker");
  } catch (const std::string& s) {
    std::printf("missing: %s", s.c_str());
  }
  std::printf("created=%d models=%zu fx=%.3f ndist=%zu\n", created, models.size(), cam.Intrinsic[0], cam.distCoeffs.size());
  return 0;
}
