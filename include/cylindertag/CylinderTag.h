// CylinderTag.h -- C++ host-side mirror of the reference's public class (header/CylinderTag.h:12-52) over the
// B200 C ABI (include/ctag.h).  Header only; link against libctag_b200.so.
//
// It keeps the reference's method names, argument meaning and error behaviour for the detection path:
//   CylinderTag(path) / CylinderTag(state)   CylinderTag.cpp:6-65   (throws std::string on file / dictionary errors)
//   detect(img, markers, adaptiveThresh, cornerSubPix, cornerSubPixDist)   CylinderTag.cpp:67-128
//       (assigns `markers` on success, leaves it untouched on the "No corner detected!" / "No feature detected!"
//        early exits and prints the same messages)
//   loadModel(path, models)                   CylinderTag.cpp:161-190
//   loadCamera(path, camera)                  CylinderTag.cpp:192-196 (OpenCV YAML 1.0, cameraMatrix + distCoeffs)
// The build image has no OpenCV C++ headers, so minimal stand-in types are defined here; define CTAG_WITH_OPENCV
// before including to get cv::Mat / cv::Point2f based overloads instead.
//   estimatePose(img, markers, models, camera, poses, useDensePoseRefine)   CylinderTag.cpp:198-209
//       (host side, ctag_estimate_pose: corner selection, undistortion, EPnP, Levenberg-Marquardt; markers without a
//        model are erased from the result like the reference's markerID == -1 poses)
//   drawAxis(img, markers, models, poses, camera, axisLength)   CylinderTag.cpp:211-246
//       (host side, ctag_draw_axis; the reference shows the overlay with imshow, here it is kept in overlay() as a
//        3-channel image -- there is no highgui in this build)
#pragma once
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../ctag.h"

#ifdef CTAG_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace ctag_api {

#ifdef CTAG_WITH_OPENCV
using Point2f = cv::Point2f;
using Point3f = cv::Point3f;
#else
struct Point2f {
  float x = 0, y = 0;
};
struct Point3f {
  float x = 0, y = 0, z = 0;
};
#endif

// 8-bit image view (what detect() needs from a cv::Mat): data, rows, cols, step, channels
struct ImageView {
  const uint8_t* data = nullptr;
  int rows = 0, cols = 0;
  size_t step = 0;
  int channels = 1;
};

// row-major int matrix standing in for cv::Mat1i
struct Mat1i {
  int rows = 0, cols = 0;
  std::vector<int32_t> data;
};

// header/corner_detector.h:16-22
struct MarkerInfo {
  int markerID = -1;
  std::vector<int> featurePos, feature_ID, feature_ID_left, feature_ID_right;
  std::vector<std::vector<Point2f>> cornerLists;
  std::vector<Point2f> feature_center;
  std::vector<float> edge_length, cr_left, cr_right;
};

// header/pose_estimation.h:12-25
struct CamInfo {
  float Intrinsic[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // row-major 3x3, dt: f
  std::vector<float> distCoeffs;                      // 5x1, dt: f
};
struct ModelInfo {
  int MarkerID = -1;
  Point3f axis, base;
  std::vector<Point3f> corners;
};
struct PoseInfo {
  int markerID = -1;
  double rvec[3] = {0, 0, 0}, tvec[3] = {0, 0, 0};
};

// CylinderTag::loadModel (CylinderTag.cpp:161-190): `model_num model_size`, then per model ID, base, axis and
// 8 * model_size rows `corner_index x y z`.
inline void load_model(const std::string& path, std::vector<ModelInfo>& reconstruct_model) {
  std::ifstream in(path);
  if (!in.is_open()) throw std::string("loadModel, could not open the model file\n");
  int model_num = 0, model_size = 0;
  in >> model_num >> model_size;
  reconstruct_model.assign(model_num, ModelInfo());
  for (int i = 0; i < model_num; ++i) {
    ModelInfo& m = reconstruct_model[i];
    in >> m.MarkerID >> m.base.x >> m.base.y >> m.base.z >> m.axis.x >> m.axis.y >> m.axis.z;
    m.corners.assign((size_t)model_size * 8, Point3f());
    for (int j = 0; j < 8 * model_size; ++j) {
      int id = 0;
      Point3f p;
      in >> id >> p.x >> p.y >> p.z;
      if (id >= 0 && id < 8 * model_size) m.corners[id] = p;
    }
  }
}

// one `data: [ ... ]` matrix of an OpenCV YAML 1.0 file
inline std::vector<float> yaml_matrix(const std::string& txt, const std::string& key) {
  std::vector<float> out;
  size_t p = txt.find(key);
  if (p == std::string::npos) return out;
  size_t a = txt.find('[', p), b = txt.find(']', a);
  if (a == std::string::npos || b == std::string::npos) return out;
  std::string body = txt.substr(a + 1, b - a - 1);
  for (char& ch : body)
    if (ch == ',' || ch == '\n') ch = ' ';
  std::stringstream ss(body);
  double v;
  while (ss >> v) out.push_back((float)v);  // dt: f
  return out;
}

// CylinderTag::loadCamera (CylinderTag.cpp:192-196)
inline void load_camera(const std::string& path, CamInfo& camera) {
  std::ifstream in(path);
  std::stringstream ss;
  ss << in.rdbuf();
  const std::string txt = ss.str();
  std::vector<float> k = yaml_matrix(txt, "cameraMatrix"), dcf = yaml_matrix(txt, "distCoeffs");
  for (size_t i = 0; i < 9 && i < k.size(); ++i) camera.Intrinsic[i] = k[i];
  camera.distCoeffs = dcf;
}

// ctag_marker view of a MarkerInfo (the fields the pose stage reads)
inline ctag_marker to_record(const MarkerInfo& m) {
  ctag_marker c = ctag_marker();
  c.marker_id = m.markerID;
  const int n = (int)m.cornerLists.size() < CTAG_MAX_FEATURES ? (int)m.cornerLists.size() : CTAG_MAX_FEATURES;
  c.n_features = n;
  for (int k = 0; k < CTAG_MAX_FEATURES; ++k) c.feature_pos[k] = -1;
  for (int k = 0; k < n; ++k) {
    c.feature_pos[k] = k < (int)m.featurePos.size() ? m.featurePos[k] : -1;
    c.feature_id[k] = k < (int)m.feature_ID.size() ? m.feature_ID[k] : -1;
    c.id_left[k] = k < (int)m.feature_ID_left.size() ? m.feature_ID_left[k] : -1;
    c.id_right[k] = k < (int)m.feature_ID_right.size() ? m.feature_ID_right[k] : -1;
    for (int q = 0; q < 8 && q < (int)m.cornerLists[k].size(); ++q)
      c.corners[k][q][0] = m.cornerLists[k][q].x, c.corners[k][q][1] = m.cornerLists[k][q].y;
  }
  return c;
}

// PoseEstimator::PnPSolver + PoseBA for every marker (pose_estimation.cpp:50-143) through ctag_estimate_pose.
// PoseInfo::markerID is the INDEX of the model in `reconstruct_model` (pose_estimation.cpp:59,69); markers whose ID has
// no model, or with too few usable corners, produce no pose (the reference erases its markerID == -1 entries,
// CylinderTag.cpp:206-208).  Host code only: no detector, no GPU.
inline void estimate_poses(const std::vector<MarkerInfo>& markers, const std::vector<ModelInfo>& reconstruct_model,
                           const CamInfo& camera, std::vector<PoseInfo>& pose) {
  pose.clear();
  for (const MarkerInfo& mk : markers) {
    int idx = -1;
    for (size_t j = 0; j < reconstruct_model.size(); ++j)
      if (reconstruct_model[j].MarkerID == mk.markerID) {
        idx = (int)j;
        break;
      }
    if (idx < 0) continue;
    const ModelInfo& model = reconstruct_model[idx];
    ctag_marker rec = to_record(mk);
    std::vector<float> pts(model.corners.size() * 3);
    for (size_t i = 0; i < model.corners.size(); ++i)
      pts[3 * i] = model.corners[i].x, pts[3 * i + 1] = model.corners[i].y, pts[3 * i + 2] = model.corners[i].z;
    PoseInfo p;
    if (ctag_estimate_pose(&rec, pts.data(), (int)model.corners.size(), camera.Intrinsic, camera.distCoeffs.data(),
                           (int)camera.distCoeffs.size(), p.rvec, p.tvec, nullptr) != CTAG_OK)
      continue;
    p.markerID = idx;
    pose.push_back(p);
  }
}

// 3-channel 8-bit image the overlay is drawn into (rows x cols x 3, interleaved, channel order of the reference's
// Scalar values)
struct Overlay {
  int rows = 0, cols = 0;
  std::vector<uint8_t> data;
};

// CylinderTag::drawAxis (CylinderTag.cpp:211-246) through ctag_gray_to_3ch / ctag_draw_axis: gray -> 3 channels, then
// per pose the projected model corners, the three arrows and the base point.  pose[i] is drawn with markers[i] like
// the reference does (so after estimatePose erased a marker without a model the pairs are shifted, SURVEY 8f-4:
// replicated, guarded against running off the end).  Host code only: no detector, no GPU.
inline void draw_axis(const ImageView& img, const std::vector<MarkerInfo>& markers, const std::vector<ModelInfo>& reconstruct_model,
                      const std::vector<PoseInfo>& pose, const CamInfo& camera, int axisLength, Overlay& out) {
  out.rows = img.rows, out.cols = img.cols;
  out.data.assign((size_t)img.rows * img.cols * 3, 0);
  int rc = ctag_gray_to_3ch(img.data, img.cols, img.rows, img.step, out.data.data(), (size_t)img.cols * 3);
  if (rc != CTAG_OK) throw std::string("drawAxis, ") + ctag_strerror(rc) + "\n";
  for (size_t i = 0; i < pose.size() && i < markers.size(); ++i) {
    const int id = pose[i].markerID;
    if (id < 0 || id >= (int)reconstruct_model.size()) continue;
    const ModelInfo& model = reconstruct_model[id];
    ctag_marker rec = to_record(markers[i]);
    std::vector<float> pts(model.corners.size() * 3);
    for (size_t k = 0; k < model.corners.size(); ++k)
      pts[3 * k] = model.corners[k].x, pts[3 * k + 1] = model.corners[k].y, pts[3 * k + 2] = model.corners[k].z;
    const float base[3] = {model.base.x, model.base.y, model.base.z}, axis[3] = {model.axis.x, model.axis.y, model.axis.z};
    rc = ctag_draw_axis(out.data.data(), img.cols, img.rows, (size_t)img.cols * 3, &rec, pts.data(), (int)model.corners.size(),
                        base, axis, camera.Intrinsic, camera.distCoeffs.data(), (int)camera.distCoeffs.size(), pose[i].rvec,
                        pose[i].tvec, axisLength);
    if (rc != CTAG_OK) throw std::string("drawAxis, ") + ctag_strerror(rc) + "\n";
  }
}

class CylinderTag {
 public:
  // Load state matrix of CylinderTag from file (CylinderTag.cpp:6-9,16-41)
  explicit CylinderTag(const std::string& path, int cuda_device = -1) {
    int rc = ctag_create_from_file(&det_, path.c_str(), cuda_device);
    if (rc == CTAG_ERR_FILE) throw std::string("load_from_file, could not open the file\n");
    if (rc == CTAG_ERR_DICTIONARY)
      throw std::string("check_dictionary, the number in state matrix must between 0 to 63\nload_from_file, illegal marker info\n");
    if (rc != CTAG_OK) throw std::string("CylinderTag, ") + ctag_strerror(rc) + ": " + ctag_last_error() + "\n";
  }
  // Manual input of the state matrix (CylinderTag.cpp:11-14,43-54).  The reference leaves featureSize unset here
  // (SURVEY C-3); it has to be given.
  CylinderTag(const Mat1i& set_state, int feature_size, int cuda_device = -1) {
    int rc = ctag_create(&det_, set_state.data.data(), set_state.rows, set_state.cols, feature_size, cuda_device);
    if (rc == CTAG_ERR_DICTIONARY)
      throw std::string("check_dictionary, the number in state matrix must between 0 to 63\nload_from_set, illegal marker info\n");
    if (rc != CTAG_OK) throw std::string("CylinderTag, ") + ctag_strerror(rc) + ": " + ctag_last_error() + "\n";
  }
  ~CylinderTag() { ctag_destroy(det_); }
  CylinderTag(const CylinderTag&) = delete;
  CylinderTag& operator=(const CylinderTag&) = delete;

  // Marker Detector (CylinderTag.cpp:67-128). `img` is 8-bit single channel.
  void detect(const ImageView& img, std::vector<MarkerInfo>& cornerList, int adaptiveThresh = 5, const bool cornerSubPix = false,
              int cornerSubPixDist = 3) {
    // the reference's cvtColor(img, imgMark, COLOR_GRAY2RGB) (CylinderTag.cpp:70) asserts on anything but one channel
    if (img.channels != 1) throw std::string("detect, the image must be 8-bit single-channel (convert BGR frames first, main.cpp:54, or use detectBatch with channels = 3)\n");
    std::vector<ctag_marker> buf(kCap);
    int n = 0, status = 0;
    int rc = ctag_detect(det_, img.data, img.cols, img.rows, img.step, adaptiveThresh, cornerSubPix ? 1 : 0, cornerSubPixDist,
                         buf.data(), kCap, &n, &status);
    if (rc != CTAG_OK) throw std::string("detect, ") + ctag_strerror(rc) + ": " + ctag_last_error() + "\n";
    if (status == CTAG_FRAME_NO_CORNER) {
      std::cout << "No corner detected!" << std::endl;  // CylinderTag.cpp:88; output untouched
      return;
    }
    if (status == CTAG_FRAME_NO_FEATURE) {
      std::cout << "No feature detected!" << std::endl;  // CylinderTag.cpp:94; output untouched
      return;
    }
    std::vector<MarkerInfo> out;
    for (int m = 0; m < n && m < kCap; ++m) out.push_back(convert(buf[m]));
    cornerList = out;  // markers_info = markers (CylinderTag.cpp:128)
  }

  // Batched detection -- an extension: the reference walks a video frame by frame (main.cpp:48-60), the B200 library
  // takes many frames per call (ctag_detect_batch).  `frames`: n images of cols x rows pixels in host memory, `step`
  // bytes per row, `frame_stride` bytes from one frame to the next; channels 1 (gray) or 3 (BGR: the caller's cvtColor
  // is done on the GPU).  out[f] is what detect() would assign for frame f, and is left EMPTY where detect() would have
  // left its argument untouched (no corner / no feature); status (optional) receives the CTAG_FRAME_* code per frame.
  void detectBatch(const uint8_t* frames, int n, int rows, int cols, size_t step, size_t frame_stride, int channels,
                   std::vector<std::vector<MarkerInfo>>& out, int adaptiveThresh = 5, const bool cornerSubPix = false,
                   int cornerSubPixDist = 3, std::vector<int>* status = nullptr) {
    std::vector<ctag_marker> buf((size_t)n * kCap);
    std::vector<int> count(n, 0);
    std::vector<ctag_frame_info> info(n);
    int rc = ctag_detect_batch(det_, frames, n, cols, rows, step, frame_stride, channels, 0, adaptiveThresh, cornerSubPix ? 1 : 0,
                               cornerSubPixDist, buf.data(), kCap, count.data(), info.data());
    if (rc != CTAG_OK) throw std::string("detectBatch, ") + ctag_strerror(rc) + ": " + ctag_last_error() + "\n";
    out.assign(n, {});
    if (status) status->assign(n, 0);
    for (int f = 0; f < n; ++f) {
      if (status) (*status)[f] = info[f].status;
      for (int m = 0; m < count[f] && m < kCap; ++m) out[f].push_back(convert(buf[(size_t)f * kCap + m]));
    }
  }

#ifdef CTAG_WITH_OPENCV
  void detect(const cv::Mat& img, std::vector<MarkerInfo>& cornerList, int adaptiveThresh = 5, const bool cornerSubPix = false,
              int cornerSubPixDist = 3) {
    detect(ImageView{img.data, img.rows, img.cols, img.step, img.channels()}, cornerList, adaptiveThresh, cornerSubPix,
           cornerSubPixDist);
  }
#endif

  // Load reconstructed model (CylinderTag.cpp:161-190)
  void loadModel(const std::string& path, std::vector<ModelInfo>& reconstruct_model) { load_model(path, reconstruct_model); }

  // Load camera intrinsic (CylinderTag.cpp:192-196): OpenCV YAML 1.0 with cameraMatrix (3x3) and distCoeffs (5x1)
  void loadCamera(const std::string& path, CamInfo& camera) { load_camera(path, camera); }

  // Marker Localization (CylinderTag.cpp:198-209).  `img` and `useDensePoseRefine` are unused, as in the reference.
  void estimatePose(const ImageView& img, const std::vector<MarkerInfo>& markers, const std::vector<ModelInfo>& reconstruct_model,
                    const CamInfo& camera, std::vector<PoseInfo>& pose, bool useDensePoseRefine = false) {
    (void)img;
    (void)useDensePoseRefine;
    estimate_poses(markers, reconstruct_model, camera, pose);
  }

  // Axis overlay (CylinderTag.cpp:211-246).  The reference opens a highgui window; here the image is kept and returned
  // by overlay() (draw_axis above does the work; host code).
  void drawAxis(const ImageView& img, const std::vector<MarkerInfo>& markers, const std::vector<ModelInfo>& reconstruct_model,
                const std::vector<PoseInfo>& pose, const CamInfo& camera, int axisLength = 5) {
    draw_axis(img, markers, reconstruct_model, pose, camera, axisLength, overlay_);
  }
  const Overlay& overlay() const { return overlay_; }

  ctag_detector* handle() { return det_; }

 private:
  static constexpr int kCap = 64;
  ctag_detector* det_ = nullptr;
  Overlay overlay_;

  static MarkerInfo convert(const ctag_marker& c) {
    MarkerInfo m;
    m.markerID = c.marker_id;
    for (int k = 0; k < c.n_features; ++k) {
      m.feature_ID.push_back(c.feature_id[k]);
      m.feature_ID_left.push_back(c.id_left[k]);
      m.feature_ID_right.push_back(c.id_right[k]);
      std::vector<Point2f> cl(8);
      for (int q = 0; q < 8; ++q) cl[q].x = c.corners[k][q][0], cl[q].y = c.corners[k][q][1];
      m.cornerLists.push_back(cl);
      Point2f ce;
      ce.x = c.center[k][0], ce.y = c.center[k][1];
      m.feature_center.push_back(ce);
      m.edge_length.push_back(c.edge_length[k]);
      m.cr_left.push_back(c.cr_left[k]);
      m.cr_right.push_back(c.cr_right[k]);
    }
    for (int k = 0; k < CTAG_MAX_FEATURES && c.feature_pos[k] >= 0 && (int)m.featurePos.size() < c.n_features; ++k)
      m.featurePos.push_back(c.feature_pos[k]);
    return m;
  }

};

}  // namespace ctag_api
