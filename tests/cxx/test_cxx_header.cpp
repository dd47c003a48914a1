// Compile-and-link check of the C++ mirror class against the shared library (no GPU needed to build; at run time it
// exercises the loaders and the error paths; detection itself is covered by the gpu tests through the same C ABI).
#include <cmath>
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
  // Pose stage (host code, no detector needed): project three features of model 0 with a known pose through the
  // pinhole model (no distortion) and recover it.
  {
    std::vector<ModelInfo> mm;
    CamInfo cc;
    load_model(argv[2], mm);
    load_camera(argv[3], cc);
    cc.distCoeffs.assign(5, 0.f);
    const double rv[3] = {0.2, -0.3, 0.1}, tv[3] = {-250.0, 100.0, 300.0};
    const double th = std::sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
    const double x = rv[0] / th, y = rv[1] / th, z = rv[2] / th, c = std::cos(th), s = std::sin(th), c1 = 1 - c;
    const double R[9] = {c + c1 * x * x, c1 * x * y - s * z, c1 * x * z + s * y, c1 * x * y + s * z, c + c1 * y * y,
                         c1 * y * z - s * x, c1 * x * z - s * y, c1 * y * z + s * x, c + c1 * z * z};
    MarkerInfo mk;
    mk.markerID = mm[0].MarkerID;
    for (int f = 3; f < 6; ++f) {
      std::vector<Point2f> cl(8);
      for (int k = 0; k < 8; ++k) {
        const Point3f& P = mm[0].corners[f * 8 + k];
        const double X = R[0] * P.x + R[1] * P.y + R[2] * P.z + tv[0], Y = R[3] * P.x + R[4] * P.y + R[5] * P.z + tv[1],
                     Z = R[6] * P.x + R[7] * P.y + R[8] * P.z + tv[2];
        cl[k].x = (float)(cc.Intrinsic[0] * X / Z + cc.Intrinsic[2]);
        cl[k].y = (float)(cc.Intrinsic[4] * Y / Z + cc.Intrinsic[5]);
      }
      mk.cornerLists.push_back(cl);
      mk.featurePos.push_back(f);
      mk.feature_ID.push_back(0);
      mk.feature_ID_left.push_back(1);
      mk.feature_ID_right.push_back(1);
    }
    MarkerInfo unknown = mk;
    unknown.markerID = 12345;  // no model: erased from the result
    std::vector<PoseInfo> poses;
    estimate_poses({unknown, mk}, mm, cc, poses);
    double er = 0, et = 0;
    if (poses.size() == 1)
      for (int k = 0; k < 3; ++k) er = std::fmax(er, std::fabs(poses[0].rvec[k] - rv[k])), et = std::fmax(et, std::fabs(poses[0].tvec[k] - tv[k]));
    std::printf("poses=%zu model_index=%d rot_err_ok=%d trans_err_ok=%d\n", poses.size(), poses.empty() ? -9 : poses[0].markerID,
                (int)(poses.size() == 1 && er < 1e-4), (int)(poses.size() == 1 && et < 3e-2));
    // drawAxis overlay (host code): a flat gray image, the recovered pose; the corner discs carry (255, 234, 32) at the
    // image points the marker was built from, everything far from the drawing stays gray
    if (poses.size() == 1) {
      const int W = 1920, H = 1200;
      std::vector<uint8_t> gray((size_t)W * H, 90);
      Overlay ov;
      // the test pose looks far off-axis: move the principal point so that the marker lands inside the image
      CamInfo cs = cc;
      const float sx = 960.f - mk.cornerLists[1][0].x, sy = 600.f - mk.cornerLists[1][0].y;
      cs.Intrinsic[2] += sx, cs.Intrinsic[5] += sy;
      draw_axis(ImageView{gray.data(), H, W, (size_t)W, 1}, {mk}, mm, poses, cs, 30, ov);
      int discs = 0, painted = 0;
      for (int f = 0; f < 3; ++f)
        for (int k = 0; k < 8; ++k) {
          if (f == 2 && k == 7) continue;  // the reference's loop bound skips the last corner
          const int px = (int)std::lround(mk.cornerLists[f][k].x + sx), py = (int)std::lround(mk.cornerLists[f][k].y + sy);
          if (px < 0 || py < 0 || px >= W || py >= H) continue;
          const uint8_t* p = &ov.data[((size_t)py * W + px) * 3];
          discs += (p[0] == 255 && p[1] == 234 && p[2] == 32);
        }
      for (size_t i = 0; i < (size_t)W * H; ++i) painted += !(ov.data[3 * i] == 90 && ov.data[3 * i + 1] == 90 && ov.data[3 * i + 2] == 90);
      std::printf("overlay=%dx%d discs=%d painted_ok=%d\n", ov.cols, ov.rows, discs, (int)(painted > 1500 && painted < 200000));
    }
  }
  return 0;
}
