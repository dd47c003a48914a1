// Generates a dictionary with generator.h, checks it, breaks it, writes it as .marker.
#include <cstdio>
#include <cstdlib>
#include "cylindertag/generator.h"
using namespace ctag_api;
int main(int argc, char** argv) {
  if (argc < 5) return 2;
  const int cols = std::atoi(argv[1]), fsz = std::atoi(argv[2]), rows = std::atoi(argv[3]);
  Mat1i cb = generate_codebook(cols, fsz, rows, 7);
  const bool ok = check_codebook(cb, fsz);
  Mat1i broken = cb;
  if (broken.rows >= 2)
    for (int c = 0; c < cols; ++c) broken.data[cols + c] = broken.data[c];  // row 1 = row 0: windows repeat
  Mat1i illegal = cb;
  if (!illegal.data.empty()) illegal.data[0] = 8 * 1 + 6;                   // digits from different halves
  std::printf("rows=%d cols=%d ok=%d broken_ok=%d illegal_ok=%d written=%d\n", cb.rows, cb.cols, (int)ok,
              (int)check_codebook(broken, fsz), (int)check_codebook(illegal, fsz), (int)write_marker_file(argv[4], cb, fsz));
  return 0;
}
