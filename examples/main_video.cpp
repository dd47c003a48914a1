// The video branch of the reference's demo (main.cpp:43-60: per frame cvtColor, clear, detect(…, 5, true, 5),
// estimatePose) on the B200 library.  There is no video decoder in this build, so the "video" is an image sequence
// (BMP/PGM/PPM files of equal size).  The frames go through the reference's frame-by-frame loop and, in one call,
// through the batched entry point the library adds (detectBatch); both must give the same markers.
//   g++ -std=c++17 -I include examples/main_video.cpp -L cylindertag_b200/lib -lctag_b200 -Wl,-rpath,$PWD/cylindertag_b200/lib
//   ./a.out CTag_2f12c.marker CTag_2f12c.model cameraParams.yml frame0.pgm frame1.pgm ...
#include <cstdio>
#include <cstring>

#include "cylindertag/CylinderTag.h"
#include "cylindertag/imageio.h"

using namespace ctag_api;

static bool same_markers(const std::vector<MarkerInfo>& a, const std::vector<MarkerInfo>& b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); ++i) {
    if (a[i].markerID != b[i].markerID || a[i].featurePos != b[i].featurePos || a[i].cornerLists.size() != b[i].cornerLists.size())
      return false;
    for (size_t j = 0; j < a[i].cornerLists.size(); ++j)
      for (int k = 0; k < 8; ++k)
        if (a[i].cornerLists[j][k].x != b[i].cornerLists[j][k].x || a[i].cornerLists[j][k].y != b[i].cornerLists[j][k].y) return false;
  }
  return true;
}

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: %s file.marker file.model cameraParams.yml frame0 [frame1 ...]\n", argv[0]);
    return 2;
  }
  try {
    CylinderTag marker(argv[1]);
    std::vector<ModelInfo> marker_model;
    CamInfo camera;
    marker.loadModel(argv[2], marker_model);
    marker.loadCamera(argv[3], camera);
    const int n = argc - 4;
    std::vector<uint8_t> all;  // the gray frames back to back, for the batched call
    int rows = 0, cols = 0;
    std::vector<std::vector<MarkerInfo>> per_frame(n);
    for (int f = 0; f < n; ++f) {
      Image frame = imread(argv[4 + f]);
      if (frame.empty()) {
        std::fprintf(stderr, "could not read %s\n", argv[4 + f]);
        return 1;
      }
      Image img_gray = bgr2gray(frame);
      if (f == 0) rows = img_gray.rows, cols = img_gray.cols;
      if (img_gray.rows != rows || img_gray.cols != cols) {
        std::fprintf(stderr, "%s: frame size differs\n", argv[4 + f]);
        return 1;
      }
      all.insert(all.end(), img_gray.data.begin(), img_gray.data.end());
      std::vector<MarkerInfo> markers;  // main.cpp:55-56: both lists start empty for every frame
      std::vector<PoseInfo> pose;
      marker.detect(img_gray.view(), markers, 5, true, 5);
      marker.estimatePose(img_gray.view(), markers, marker_model, camera, pose, false);
      per_frame[f] = markers;
      std::printf("frame %d markers %zu poses %zu ids", f, markers.size(), pose.size());
      for (const MarkerInfo& m : markers) std::printf(" %d", m.markerID);
      std::printf("\n");
      for (const PoseInfo& p : pose)
        std::printf("frame %d pose model %d rvec %.5f %.5f %.5f tvec %.4f %.4f %.4f\n", f, p.markerID, p.rvec[0], p.rvec[1],
                    p.rvec[2], p.tvec[0], p.tvec[1], p.tvec[2]);
    }
    std::vector<std::vector<MarkerInfo>> batched;
    marker.detectBatch(all.data(), n, rows, cols, (size_t)cols, (size_t)rows * cols, 1, batched, 5, true, 5);
    int equal = 0;
    for (int f = 0; f < n; ++f) equal += same_markers(per_frame[f], batched[f]);
    std::printf("batched call equals the frame loop on %d of %d frames\n", equal, n);
  } catch (const std::string& s) {
    std::fprintf(stderr, "%s", s.c_str());
    return 1;
  }
  return 0;
}
