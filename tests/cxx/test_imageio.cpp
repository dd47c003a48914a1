// Reads an image with ctag_api::imread and prints shape + FNV-1a of the (gray-converted) pixels.
#include <cstdio>
#include "cylindertag/imageio.h"
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  ctag_api::Image im = ctag_api::imread(argv[1]);
  if (im.empty()) {
    std::printf("empty\n");
    return 0;
  }
  ctag_api::Image g = ctag_api::bgr2gray(im);
  unsigned long long h = 1469598103934665603ull;
  for (unsigned char v : g.data) h = (h ^ v) * 1099511628211ull;
  std::printf("%d %d %d %llu\n", im.rows, im.cols, im.channels, h);
  return 0;
}
