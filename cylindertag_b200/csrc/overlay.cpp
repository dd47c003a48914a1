// Host-side drawAxis overlay: CylinderTag::drawAxis (CylinderTag.cpp:211-246) without OpenCV and without
// the highgui window.  Plain C++ (no CUDA): a debugging aid, not part of the detection path.
//
//   1. cv::projectPoints with the 5-coefficient lens model                     CylinderTag.cpp:234
//   2. filled circles on every model corner of the decoded features (r = 5)    :235-237
//   3. three arrowed lines from the projected base point (thickness 10, tip length 0.2): the model axis and the two
//      fixed directions the reference hard-codes                               :229-231, 239-241
//   4. a filled circle on the base point (r = 8)                               :242
//
// Rasterisation: discs are the pixels within the radius; thick segments are drawn by their distance field with a
// one-pixel linear edge ramp calibrated on OpenCV's LINE_AA strokes (full coverage up to thickness/2 + 0.25).  The
// result agrees with OpenCV's output up to edge pixels; nothing downstream reads it.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/ctag.h"

namespace {

void rodrigues(const double* r, double* R) {
  const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (th < 1e-300) {
    R[0] = R[4] = R[8] = 1;
    R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0;
    return;
  }
  const double x = r[0] / th, y = r[1] / th, z = r[2] / th, c = cos(th), s = sin(th), c1 = 1 - c;
  R[0] = c + c1 * x * x, R[1] = c1 * x * y - s * z, R[2] = c1 * x * z + s * y;
  R[3] = c1 * x * y + s * z, R[4] = c + c1 * y * y, R[5] = c1 * y * z - s * x;
  R[6] = c1 * x * z - s * y, R[7] = c1 * y * z + s * x, R[8] = c + c1 * z * z;
}

struct Canvas {
  uint8_t* px;
  int w, h;
  size_t pitch;
  // blend colour c into pixel (x, y) with coverage a in [0, 1]
  void blend(int x, int y, const double* c, double a) const {
    if (x < 0 || y < 0 || x >= w || y >= h || a <= 0) return;
    uint8_t* p = px + (size_t)y * pitch + 3 * (size_t)x;
    if (a >= 1) {
      for (int k = 0; k < 3; ++k) p[k] = (uint8_t)c[k];
      return;
    }
    for (int k = 0; k < 3; ++k) p[k] = (uint8_t)lrint(p[k] + a * (c[k] - p[k]));
  }
};

// every pixel whose centre is within `radius` of the segment a-b (a == b: a disc); `aa`: one-pixel edge ramp
void draw_capsule(const Canvas& cv, double ax, double ay, double bx, double by, double radius, const double* colour, bool aa) {
  if (!(isfinite(ax) && isfinite(ay) && isfinite(bx) && isfinite(by))) return;
  const double pad = radius + 1;
  const double fx0 = std::min(ax, bx) - pad, fx1 = std::max(ax, bx) + pad, fy0 = std::min(ay, by) - pad, fy1 = std::max(ay, by) + pad;
  if (fx1 < 0 || fy1 < 0 || fx0 > cv.w || fy0 > cv.h) return;
  const int x0 = (int)std::max(0.0, floor(fx0)), x1 = (int)std::min((double)cv.w - 1, ceil(fx1));
  const int y0 = (int)std::max(0.0, floor(fy0)), y1 = (int)std::min((double)cv.h - 1, ceil(fy1));
  const double dx = bx - ax, dy = by - ay, len2 = dx * dx + dy * dy;
  for (int y = y0; y <= y1; ++y)
    for (int x = x0; x <= x1; ++x) {
      double t = len2 > 0 ? ((x - ax) * dx + (y - ay) * dy) / len2 : 0.0;
      t = t < 0 ? 0 : (t > 1 ? 1 : t);
      const double ex = x - (ax + t * dx), ey = y - (ay + t * dy);
      const double d = sqrt(ex * ex + ey * ey);
      const double a = aa ? radius + 1.25 - d : (d <= radius ? 1.0 : 0.0);  // OpenCV's thick LINE_AA strokes: full to r + 0.25
      cv.blend(x, y, colour, a);
    }
}

// cv::arrowedLine(img, p, q, colour, thickness, LINE_AA, 0, tipLength): shaft plus two tip strokes at +-45 degrees of
// length tipLength * |pq|
void draw_arrow(const Canvas& cv, double px, double py, double qx, double qy, const double* colour, double thickness, double tip) {
  draw_capsule(cv, px, py, qx, qy, 0.5 * thickness, colour, true);
  const double size = sqrt((px - qx) * (px - qx) + (py - qy) * (py - qy)) * tip;
  const double ang = atan2(py - qy, px - qx);
  const double kPi4 = 0.78539816339744830962;
  for (int s = -1; s <= 1; s += 2) {
    const double tx = lrint(qx + size * cos(ang + s * kPi4)), ty = lrint(qy + size * sin(ang + s * kPi4));
    draw_capsule(cv, tx, ty, qx, qy, 0.5 * thickness, colour, true);
  }
}

}  // namespace

extern "C" {

// cv::projectPoints (calib3d): x_cam = R(rvec) X + t, pinhole division, radial (k1, k2, k3) and tangential (p1, p2)
// distortion, then the camera matrix.  dist order k1 k2 p1 p2 k3 as in cameraParams.yml.
int ctag_project_points(const float* points3, int n, const double* rvec, const double* tvec, const float* intrinsic,
                        const float* dist, int n_dist, float* out_xy) {
  if (n < 0 || (n > 0 && (!points3 || !out_xy)) || !rvec || !tvec || !intrinsic || (n_dist > 0 && !dist)) return CTAG_ERR_ARG;
  double R[9];
  rodrigues(rvec, R);
  const double fx = intrinsic[0], fy = intrinsic[4], cx = intrinsic[2], cy = intrinsic[5];
  const double k1 = n_dist > 0 ? dist[0] : 0, k2 = n_dist > 1 ? dist[1] : 0, p1 = n_dist > 2 ? dist[2] : 0,
               p2 = n_dist > 3 ? dist[3] : 0, k3 = n_dist > 4 ? dist[4] : 0;
  for (int i = 0; i < n; ++i) {
    const double X = points3[3 * i], Y = points3[3 * i + 1], Z = points3[3 * i + 2];
    const double xc = R[0] * X + R[1] * Y + R[2] * Z + tvec[0];
    const double yc = R[3] * X + R[4] * Y + R[5] * Z + tvec[1];
    double zc = R[6] * X + R[7] * Y + R[8] * Z + tvec[2];
    zc = zc ? 1.0 / zc : 1.0;
    const double x = xc * zc, y = yc * zc;
    const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
    const double cdist = 1 + k1 * r2 + k2 * r4 + k3 * r6;
    const double xd = x * cdist + p1 * a1 + p2 * a2, yd = y * cdist + p1 * a3 + p2 * a1;
    out_xy[2 * i] = (float)(xd * fx + cx);
    out_xy[2 * i + 1] = (float)(yd * fy + cy);
  }
  return CTAG_OK;
}

// cvtColor(img, imgMark, COLOR_GRAY2RGB) (CylinderTag.cpp:214): the gray value on all three channels
int ctag_gray_to_3ch(const uint8_t* gray, int w, int h, size_t pitch, uint8_t* out3, size_t out_pitch) {
  if (!gray || !out3 || w <= 0 || h <= 0 || pitch < (size_t)w || out_pitch < 3 * (size_t)w) return CTAG_ERR_ARG;
  for (int y = 0; y < h; ++y) {
    const uint8_t* s = gray + (size_t)y * pitch;
    uint8_t* d = out3 + (size_t)y * out_pitch;
    for (int x = 0; x < w; ++x) d[3 * x] = d[3 * x + 1] = d[3 * x + 2] = s[x];
  }
  return CTAG_OK;
}

// The body of drawAxis's loop over poses (CylinderTag.cpp:219-243) for one (marker, pose) pair.
int ctag_draw_axis(uint8_t* img3, int w, int h, size_t pitch, const ctag_marker* mk, const float* model_corners,
                   int n_model_corners, const float* base, const float* axis, const float* intrinsic, const float* dist,
                   int n_dist, const double* rvec, const double* tvec, int axis_length) {
  if (!img3 || w <= 0 || h <= 0 || pitch < 3 * (size_t)w || !mk || !model_corners || !base || !axis) return CTAG_ERR_ARG;
  const int nf = std::min(std::max(mk->n_features, 0), (int)CTAG_MAX_FEATURES);
  std::vector<float> pts;
  pts.reserve((size_t)3 * (8 * nf + 4));
  for (int j = 0; j < nf; ++j)
    for (int k = 0; k < 8; ++k) {
      const int id = mk->feature_pos[j] * 8 + k;
      if (id < 0 || id >= n_model_corners) return CTAG_ERR_ARG;
      for (int c = 0; c < 3; ++c) pts.push_back(model_corners[3 * id + c]);
    }
  // base, base + axis * L, base + (0.0372, 0.0372, 0.9986) * L, base + (0.9980, -0.0520, -0.0353) * L  (float arithmetic
  // of Point3f, CylinderTag.cpp:228-231)
  const float L = (float)axis_length;
  const float fixed[2][3] = {{0.0372f, 0.0372f, 0.9986f}, {0.9980f, -0.0520f, -0.0353f}};
  for (int c = 0; c < 3; ++c) pts.push_back(base[c]);
  for (int c = 0; c < 3; ++c) pts.push_back(base[c] + axis[c] * L);
  for (int d = 0; d < 2; ++d)
    for (int c = 0; c < 3; ++c) pts.push_back(base[c] + fixed[d][c] * L);
  const int n = (int)pts.size() / 3;
  std::vector<float> xy((size_t)2 * n);
  const int rc = ctag_project_points(pts.data(), n, rvec, tvec, intrinsic, dist, n_dist, xy.data());
  if (rc != CTAG_OK) return rc;
  const Canvas cv{img3, w, h, pitch};
  // Point2f -> Point rounds (cvRound)
  auto px = [&](int i, double& x, double& y) { x = (double)lrint(xy[2 * i]), y = (double)lrint(xy[2 * i + 1]); };
  const double corner_col[3] = {255, 234, 32}, base_col[3] = {247, 235, 235};
  const double arrow_col[3][3] = {{255, 0, 0}, {0, 255, 0}, {0, 0, 255}};
  double x, y, ox, oy;
  // the reference's loop bound is size() - 5: the last model corner is not marked (:235)
  for (int i = 0; i < n - 5; ++i) {
    px(i, x, y);
    draw_capsule(cv, x, y, x, y, 5, corner_col, false);
  }
  px(n - 4, ox, oy);
  for (int a = 0; a < 3; ++a) {
    px(n - 3 + a, x, y);
    draw_arrow(cv, ox, oy, x, y, arrow_col[a], 10, 0.2);
  }
  draw_capsule(cv, ox, oy, ox, oy, 8, base_col, false);
  return CTAG_OK;
}

}  // extern "C"
