// The image branch of the reference's demo (main.cpp:27-41: imread, CylinderTag(marker), loadModel, loadCamera,
// cvtColor, detect(…, 5, true, 5), estimatePose, drawAxis(…, 30)) on the B200 library, without OpenCV or Ceres.
//   g++ -std=c++17 -I include examples/main_image.cpp -L cylindertag_b200/lib -lctag_b200 -Wl,-rpath,$PWD/cylindertag_b200/lib
//   ./a.out test.bmp CTag_2f12c.marker CTag_2f12c.model cameraParams.yml [overlay.ppm]
#include <cstdio>

#include "cylindertag/CylinderTag.h"
#include "cylindertag/imageio.h"

using namespace ctag_api;

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: %s image.bmp|.pgm|.ppm file.marker file.model cameraParams.yml [overlay.ppm]\n", argv[0]);
    return 2;
  }
  try {
    Image frame = imread(argv[1]);
    if (frame.empty()) {
      std::fprintf(stderr, "could not read %s\n", argv[1]);
      return 1;
    }
    CylinderTag marker(argv[2]);
    std::vector<ModelInfo> marker_model;
    CamInfo camera;
    marker.loadModel(argv[3], marker_model);
    marker.loadCamera(argv[4], camera);
    Image img_gray = bgr2gray(frame);
    std::vector<MarkerInfo> markers;
    std::vector<PoseInfo> pose;
    marker.detect(img_gray.view(), markers, 5, true, 5);
    marker.estimatePose(img_gray.view(), markers, marker_model, camera, pose, false);
    std::printf("markers %zu poses %zu\n", markers.size(), pose.size());
    for (const MarkerInfo& m : markers) {
      std::printf("marker %d features %zu pos", m.markerID, m.cornerLists.size());
      for (int p : m.featurePos) std::printf(" %d", p);
      std::printf("\n");
    }
    for (const PoseInfo& p : pose)
      std::printf("pose model %d id %d rvec %.5f %.5f %.5f tvec %.4f %.4f %.4f\n", p.markerID, marker_model[p.markerID].MarkerID,
                  p.rvec[0], p.rvec[1], p.rvec[2], p.tvec[0], p.tvec[1], p.tvec[2]);
    if (argc > 5) {  // main.cpp:41 shows the overlay in a window; here it goes to a file
      marker.drawAxis(img_gray.view(), markers, marker_model, pose, camera, 30);
      const Overlay& ov = marker.overlay();
      if (!imwrite_pnm(argv[5], ov.data.data(), ov.rows, ov.cols, 3)) {
        std::fprintf(stderr, "could not write %s\n", argv[5]);
        return 1;
      }
    }
  } catch (const std::string& s) {
    std::fprintf(stderr, "%s", s.c_str());
    return 1;
  }
  return 0;
}
