// TEST INFRASTRUCTURE ONLY -- never linked into the product (cylindertag_b200/).
//
// C ABI around the UNMODIFIED reference class (header/CylinderTag.h:12-52), so that pytest / bench.py can drive
// /root/reference's own CylinderTag::detect (CylinderTag.cpp:67-128), every corner_detector stage behind it
// (corner_detector.cpp:28-1324) and estimatePose (CylinderTag.cpp:198-209, pose_estimation.cpp:50-143) through ctypes.
// The reference sources are compiled where they lie (oracle/build_ref.py); nothing of them is copied here.
//
// Stage dumps: the reference keeps its intermediates in private members (quadAreas_labeled, corners, features;
// corner_detector::img_labeled / stats).  This translation unit includes the reference headers with `private`
// re-spelled as `public` (access only; no layout change with GCC) and reads those members after detect() returns.
//
// Oracle decisions where the reference has undefined behaviour:
//  C-1  the shim's Mat zero-fills, so the threshold's unwritten border tiles are 0;
//  C-2  corner_detector::ID_left / ID_right have no initialiser and carry over between frames: ref_detect() sets both
//       to 0 before every call (reset_ids != 0);
//  C-4  isVisited[1000] / father[100] / code[20]: a frame with > 1000 quads or > 100 features overruns them inside the
//       reference; ref_detect() reports such frames as flagged (the call has already happened; results are void).
#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include <pthread.h>

#include "ctag.h"  // the product's public C header: only for the POD ctag_marker layout the batch leg fills

#include <opencv2/core.hpp>
#include "ceres/ceres.h"
#include "ceres/rotation.h"
#include "glog/logging.h"

#define private public
#define protected public
#include "header/CylinderTag.h"
#undef private
#undef protected

namespace {

struct RefHandle {
    CylinderTag* tag = nullptr;
    std::vector<MarkerInfo> markers;
    bool untouched = true;  // detect() left the output vector alone (early exits, CylinderTag.cpp:87-96)
    int status = 0;         // 0 ok, 1 "No corner detected!", 2 "No feature detected!"
    int flagged = 0;
    int half_w = 0, half_h = 0;
    std::vector<ModelInfo> models;
    CamInfo camera;
    std::vector<PoseInfo> poses;
    std::string error;
    bool quiet = true;
};

const int SENTINEL_ID = -777;

// discards what the reference prints ("No corner detected!", CylinderTag.cpp:88,94); stateless, so safe from any thread
struct NullBuf : std::streambuf {
    int overflow(int c) override { return traits_type::not_eof(c); }
    std::streamsize xsputn(const char*, std::streamsize n) override { return n; }
};
NullBuf g_nullbuf;

void set_err(char* err, int cap, const std::string& s) {
    if (err && cap > 0) { std::strncpy(err, s.c_str(), cap - 1); err[cap - 1] = 0; }
}

// runs fn on a thread with a big stack: get_orientedEdgePoints (corner_detector.cpp:407-418) recurses once per
// boundary pixel
struct BigStackCall { void (*fn)(void*); void* arg; };
void* big_stack_tramp(void* p) { BigStackCall* c = (BigStackCall*)p; c->fn(c->arg); return nullptr; }
void run_big_stack(void (*fn)(void*), void* arg) {
    pthread_attr_t a;
    pthread_attr_init(&a);
    pthread_attr_setstacksize(&a, (size_t)256 << 20);
    BigStackCall c = {fn, arg};
    pthread_t t;
    if (pthread_create(&t, &a, big_stack_tramp, &c) != 0) { fn(arg); pthread_attr_destroy(&a); return; }
    pthread_join(t, nullptr);
    pthread_attr_destroy(&a);
}

struct DetectArgs {
    RefHandle* h; const unsigned char* gray; int w, hgt; size_t pitch; int win, subpix, dist, reset_ids; int rc;
};

void detect_body(void* p) {
    DetectArgs* a = (DetectArgs*)p;
    RefHandle* h = a->h;
    std::streambuf* old = nullptr;
    if (h->quiet) old = std::cout.rdbuf(&g_nullbuf);
    try {
        cv::Mat img(a->hgt, a->w, CV_8UC1, (void*)a->gray, a->pitch);
        if (a->reset_ids) { h->tag->detector.ID_left = 0; h->tag->detector.ID_right = 0; }
        MarkerInfo s; s.markerID = SENTINEL_ID;
        h->markers.assign(1, s);
        h->tag->detect(img, h->markers, a->win, a->subpix != 0, a->dist);
        h->untouched = h->markers.size() == 1 && h->markers[0].markerID == SENTINEL_ID;
        if (h->untouched) h->markers.clear();
        h->status = !h->untouched ? 0 : h->tag->corners.empty() ? 1 : 2;
        h->flagged = h->tag->corners.size() > 1000 || h->tag->features.size() > 100;
        h->half_w = a->w / 2; h->half_h = a->hgt / 2;
        a->rc = (int)h->markers.size();
    } catch (const std::string& s) { h->error = s; a->rc = -1; }
    catch (const std::exception& e) { h->error = e.what(); a->rc = -1; }
    catch (...) { h->error = "unknown exception"; a->rc = -1; }
    if (old) std::cout.rdbuf(old);
}

}  // namespace

extern "C" {

__attribute__((visibility("default"))) void* ref_create_from_file(const char* marker_path, char* err, int errcap) {
    RefHandle* h = new RefHandle();
    try { h->tag = new CylinderTag(std::string(marker_path)); }
    catch (const std::string& s) { set_err(err, errcap, s); delete h; return nullptr; }
    catch (const std::exception& e) { set_err(err, errcap, e.what()); delete h; return nullptr; }
    return h;
}

// CylinderTag(const Mat1i&) never sets featureSize (C-3): it is written here.
__attribute__((visibility("default"))) void* ref_create_from_state(const int* state, int rows, int cols, int feature_size, char* err, int errcap) {
    RefHandle* h = new RefHandle();
    try {
        cv::Mat1i st(rows, cols);
        std::memcpy(st.data, state, sizeof(int) * (size_t)rows * cols);
        h->tag = new CylinderTag(st);
        h->tag->featureSize = feature_size;
    }
    catch (const std::string& s) { set_err(err, errcap, s); delete h; return nullptr; }
    catch (const std::exception& e) { set_err(err, errcap, e.what()); delete h; return nullptr; }
    return h;
}

__attribute__((visibility("default"))) void ref_destroy(void* hp) {
    RefHandle* h = (RefHandle*)hp;
    if (!h) return;
    delete h->tag;
    delete h;
}

__attribute__((visibility("default"))) void ref_set_quiet(void* hp, int quiet) { ((RefHandle*)hp)->quiet = quiet != 0; }
__attribute__((visibility("default"))) const char* ref_last_error(void* hp) { return ((RefHandle*)hp)->error.c_str(); }
__attribute__((visibility("default"))) int ref_feature_size(void* hp) { return ((RefHandle*)hp)->tag->featureSize; }
__attribute__((visibility("default"))) void ref_dictionary_shape(void* hp, int* rows, int* cols) {
    RefHandle* h = (RefHandle*)hp; *rows = h->tag->state.rows; *cols = h->tag->state.cols;
}
__attribute__((visibility("default"))) void ref_dictionary(void* hp, int* out) {
    RefHandle* h = (RefHandle*)hp;
    for (int i = 0; i < h->tag->state.rows; i++) for (int j = 0; j < h->tag->state.cols; j++) *out++ = h->tag->state.at<int>(i, j);
}

// CylinderTag::detect on one 8-bit gray frame.  Returns the marker count, or -1 (ref_last_error).
__attribute__((visibility("default"))) int ref_detect(void* hp, const unsigned char* gray, int w, int hgt, size_t pitch, int adaptive_thresh, int subpix,
                                                      int subpix_dist, int reset_ids) {
    DetectArgs a = {(RefHandle*)hp, gray, w, hgt, pitch, adaptive_thresh, subpix, subpix_dist, reset_ids, 0};
    run_big_stack(detect_body, &a);
    return a.rc;
}

// counts[9]: n_labels (incl. background), n_legal, n_quads, n_features, n_markers, status, flagged, total features in
// markers, n_groups (markerOrganization's cnt)
__attribute__((visibility("default"))) void ref_counts(void* hp, int* c) {
    RefHandle* h = (RefHandle*)hp;
    c[0] = h->tag->detector.nccomp_area;
    c[1] = (int)h->tag->quadAreas_labeled.size();
    c[2] = (int)h->tag->corners.size();
    c[3] = (int)h->tag->features.size();
    c[4] = (int)h->markers.size();
    c[5] = h->status;
    c[6] = h->flagged;
    int t = 0;
    for (const auto& m : h->markers) t += (int)m.cornerLists.size();
    c[7] = t;
    c[8] = h->untouched ? 0 : h->tag->detector.cnt;
}

// label image of connectedComponentLabeling (corner_detector.cpp:82), half_h x half_w int32
__attribute__((visibility("default"))) int ref_labels(void* hp, int* out) {
    RefHandle* h = (RefHandle*)hp;
    const cv::Mat& L = h->tag->detector.img_labeled;
    if (L.empty()) return 0;
    for (int i = 0; i < L.rows; i++) std::memcpy(out + (size_t)i * L.cols, L.ptr<int>(i), sizeof(int) * L.cols);
    return L.rows * L.cols;
}

// legal components in list order: area, x0, y0, x1, y1 (5 ints each), from the pixel lists the reference built
__attribute__((visibility("default"))) int ref_components(void* hp, int* out, int cap) {
    RefHandle* h = (RefHandle*)hp;
    int n = 0;
    for (const auto& comp : h->tag->quadAreas_labeled) {
        if (n >= cap) break;
        int x0 = INT_MAX, y0 = INT_MAX, x1 = INT_MIN, y1 = INT_MIN;
        for (const auto& p : comp) { x0 = std::min(x0, p.x); x1 = std::max(x1, p.x); y0 = std::min(y0, p.y); y1 = std::max(y1, p.y); }
        int* o = out + 5 * n++;
        o[0] = (int)comp.size(); o[1] = x0; o[2] = y0; o[3] = x1; o[4] = y1;
    }
    return n;
}

__attribute__((visibility("default"))) int ref_quads(void* hp, float* out, int cap) {
    RefHandle* h = (RefHandle*)hp;
    int n = 0;
    for (const auto& q : h->tag->corners) {
        if (n >= cap) break;
        for (int k = 0; k < 4; k++) { out[8 * n + 2 * k] = q[k].x; out[8 * n + 2 * k + 1] = q[k].y; }
        n++;
    }
    return n;
}

// features as detect() left them (refined when cornerSubPix): corners 8x2, centre 2, angle 1 per feature
__attribute__((visibility("default"))) int ref_features(void* hp, float* corners, float* center, float* angle, int cap) {
    RefHandle* h = (RefHandle*)hp;
    int n = 0;
    for (const auto& f : h->tag->features) {
        if (n >= cap) break;
        for (int k = 0; k < 8; k++) { corners[16 * n + 2 * k] = f.corners[k].x; corners[16 * n + 2 * k + 1] = f.corners[k].y; }
        center[2 * n] = f.feature_center.x; center[2 * n + 1] = f.feature_center.y;
        angle[n] = f.feature_angle;
        n++;
    }
    return n;
}

// per marker: id, n_features, n_pos (3 ints each)
__attribute__((visibility("default"))) int ref_marker_summary(void* hp, int* out, int cap) {
    RefHandle* h = (RefHandle*)hp;
    int n = 0;
    for (const auto& m : h->markers) {
        if (n >= cap) break;
        out[3 * n] = m.markerID; out[3 * n + 1] = (int)m.cornerLists.size(); out[3 * n + 2] = (int)m.featurePos.size();
        n++;
    }
    return n;
}

// concatenated per-feature data of all markers, sized by the totals of ref_marker_summary
__attribute__((visibility("default"))) void ref_marker_data(void* hp, int* feature_pos, int* feature_id, int* id_left, int* id_right, float* cr_left,
                                                            float* cr_right, float* edge_length, float* center, float* corners) {
    RefHandle* h = (RefHandle*)hp;
    for (const auto& m : h->markers) {
        for (int v : m.featurePos) *feature_pos++ = v;
        for (int v : m.feature_ID) *feature_id++ = v;
        for (int v : m.feature_ID_left) *id_left++ = v;
        for (int v : m.feature_ID_right) *id_right++ = v;
        for (float v : m.cr_left) *cr_left++ = v;
        for (float v : m.cr_right) *cr_right++ = v;
        for (float v : m.edge_length) *edge_length++ = v;
        for (const auto& c : m.feature_center) { *center++ = c.x; *center++ = c.y; }
        for (const auto& f : m.cornerLists) for (int k = 0; k < 8; k++) { *corners++ = f[k].x; *corners++ = f[k].y; }
    }
}

// loadModel + loadCamera (CylinderTag.cpp:161-196).  0 = ok.
__attribute__((visibility("default"))) int ref_load_model_camera(void* hp, const char* model_path, const char* camera_path) {
    RefHandle* h = (RefHandle*)hp;
    try {
        h->models.clear();
        h->tag->loadModel(std::string(model_path), h->models);
        h->tag->loadCamera(std::string(camera_path), h->camera);
        if (h->camera.Intrinsic.empty()) { h->error = "camera file unreadable"; return -1; }
    } catch (const std::string& s) { h->error = s; return -1; }
    catch (const std::exception& e) { h->error = e.what(); return -1; }
    return 0;
}

// estimatePose on the markers of the last ref_detect (CylinderTag.cpp:198-209).  Returns the pose count (after the
// reference erased poses without a model); ids[i] = index into the model list; rt = rvec(3) tvec(3) doubles per pose.
__attribute__((visibility("default"))) int ref_estimate_pose(void* hp, int* ids, double* rt, int cap) {
    RefHandle* h = (RefHandle*)hp;
    try {
        cv::Mat dummy;
        h->poses.clear();
        h->tag->estimatePose(dummy, h->markers, h->models, h->camera, h->poses, false);
    } catch (const std::string& s) { h->error = s; return -1; }
    catch (const std::exception& e) { h->error = e.what(); return -1; }
    int n = 0;
    for (const auto& p : h->poses) {
        if (n >= cap) break;
        ids[n] = p.markerID;
        for (int k = 0; k < 3; k++) { rt[6 * n + k] = p.rvec.at<double>(k, 0); rt[6 * n + 3 + k] = p.tvec.at<double>(k, 0); }
        n++;
    }
    return n;
}

// ---- timing leg: the demo's per-frame loop (main.cpp:52-58 without pose / display) over a batch, one CylinderTag per
// host thread, frames dealt round-robin.  channels 3 = BGR (cvtColor as in main.cpp:54), 1 = gray.
struct BatchArgs {
    const int* state; int rows, cols, fs;
    const unsigned char* frames; int n, w, h, channels, win, subpix, dist, tid, nthreads;
    int* marker_counts; int* marker_ids; int ids_cap; int rc;
    int* counts8; ctag_marker* records; int rec_cap;
};

// MarkerInfo -> POD record (features beyond CTAG_MAX_FEATURES are dropped; n_features keeps the true count)
static void fill_record(const MarkerInfo& m, int frame, ctag_marker* r) {
    std::memset(r, 0, sizeof(*r));
    r->marker_id = m.markerID;
    r->n_features = (int)m.cornerLists.size();
    r->inverse = -1;  // not recorded by the reference; oracle/ref_api.py derives it from the dictionary
    r->frame = frame;
    for (int k = 0; k < CTAG_MAX_FEATURES; k++) r->feature_pos[k] = -1;
    for (size_t k = 0; k < m.featurePos.size() && k < CTAG_MAX_FEATURES; k++) r->feature_pos[k] = m.featurePos[k];
    for (size_t k = 0; k < m.cornerLists.size() && k < CTAG_MAX_FEATURES; k++) {
        r->feature_id[k] = m.feature_ID[k];
        r->id_left[k] = m.feature_ID_left[k];
        r->id_right[k] = m.feature_ID_right[k];
        r->cr_left[k] = m.cr_left[k];
        r->cr_right[k] = m.cr_right[k];
        r->edge_length[k] = m.edge_length[k];
        r->center[k][0] = m.feature_center[k].x; r->center[k][1] = m.feature_center[k].y;
        for (int c = 0; c < 8; c++) { r->corners[k][c][0] = m.cornerLists[k][c].x; r->corners[k][c][1] = m.cornerLists[k][c].y; }
    }
}

static void* batch_worker(void* p) {
    BatchArgs* a = (BatchArgs*)p;
    try {
        cv::Mat1i st(a->rows, a->cols);
        std::memcpy(st.data, a->state, sizeof(int) * (size_t)a->rows * a->cols);
        CylinderTag tag(st);
        tag.featureSize = a->fs;
        const size_t fbytes = (size_t)a->w * a->h * a->channels;
        std::vector<MarkerInfo> markers;
        for (int f = a->tid; f < a->n; f += a->nthreads) {
            cv::Mat frame(a->h, a->w, a->channels == 3 ? CV_8UC3 : CV_8UC1, (void*)(a->frames + fbytes * f)), gray;
            if (a->channels == 3) cv::cvtColor(frame, gray, cv::COLOR_BGR2GRAY); else gray = frame;
            markers.clear();
            tag.detector.ID_left = 0; tag.detector.ID_right = 0;
            MarkerInfo sentinel; sentinel.markerID = SENTINEL_ID;
            markers.assign(1, sentinel);
            tag.detect(gray, markers, a->win, a->subpix != 0, a->dist);
            const bool untouched = markers.size() == 1 && markers[0].markerID == SENTINEL_ID;
            if (untouched) markers.clear();
            if (a->counts8) {
                int* c = a->counts8 + 8 * (size_t)f;
                c[0] = tag.detector.nccomp_area; c[1] = (int)tag.quadAreas_labeled.size(); c[2] = (int)tag.corners.size();
                c[3] = (int)tag.features.size(); c[4] = untouched ? 0 : tag.detector.cnt; c[5] = (int)markers.size();
                c[6] = !untouched ? 0 : tag.corners.empty() ? 1 : 2;
                c[7] = tag.corners.size() > 1000 || tag.features.size() > 100;
            }
            if (a->records) for (int k = 0; k < (int)markers.size() && k < a->rec_cap; k++) fill_record(markers[k], f, a->records + (size_t)f * a->rec_cap + k);
            a->marker_counts[f] = (int)markers.size();
            if (a->marker_ids) for (int k = 0; k < a->ids_cap; k++) a->marker_ids[(size_t)f * a->ids_cap + k] = k < (int)markers.size() ? markers[k].markerID : -1;
        }
    } catch (...) { a->rc = -1; }
    return nullptr;
}

__attribute__((visibility("default"))) int ref_detect_batch_mt(const int* state, int rows, int cols, int fs, const unsigned char* frames, int n, int w, int h,
                                                               int channels, int adaptive_thresh, int subpix, int subpix_dist, int threads,
                                                               int* marker_counts, int* marker_ids, int ids_cap, int* counts8,
                                                               void* records, int rec_cap) {
    if (threads < 1) threads = 1;
    if (threads > n) threads = n > 0 ? n : 1;
    // "No corner detected!" goes to cout from every thread; silence the shared stream for the duration of the batch
    std::streambuf* old = std::cout.rdbuf(&g_nullbuf);
    std::vector<BatchArgs> args(threads);
    std::vector<pthread_t> tids(threads);
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, (size_t)256 << 20);
    for (int t = 0; t < threads; t++) {
        args[t] = BatchArgs{state, rows, cols, fs, frames, n, w, h, channels, adaptive_thresh, subpix, subpix_dist, t, threads, marker_counts, marker_ids, ids_cap, 0,
                            counts8, (ctag_marker*)records, rec_cap};
        pthread_create(&tids[t], &at, batch_worker, &args[t]);
    }
    int rc = 0;
    for (int t = 0; t < threads; t++) { pthread_join(tids[t], nullptr); if (args[t].rc) rc = -1; }
    pthread_attr_destroy(&at);
    std::cout.rdbuf(old);
    return rc;
}

// ---- direct access to the stand-in's primitives, for the known-answer tests against cv2 (tests/test_ref_pinning.py) --
__attribute__((visibility("default"))) int shim_resize_cubic_u8(const unsigned char* src, int sw, int sh, unsigned char* dst, int dw, int dh) {
    try {
        cv::Mat s(sh, sw, CV_8UC1, (void*)src), d;
        cv::resize(s, d, cv::Size(dw, dh), 0.5, 0.5, cv::INTER_CUBIC);
        for (int i = 0; i < dh; i++) std::memcpy(dst + (size_t)i * dw, d.ptr<unsigned char>(i), dw);
    } catch (...) { return -1; }
    return 0;
}
__attribute__((visibility("default"))) int shim_ccl(const unsigned char* img, int w, int h, int* labels, int* stats, int stats_cap) {
    try {
        cv::Mat s(h, w, CV_8UC1, (void*)img), lab, st, cen;
        int n = cv::connectedComponentsWithStats(s, lab, st, cen, 8, CV_32S, cv::CCL_BBDT);
        for (int i = 0; i < h; i++) std::memcpy(labels + (size_t)i * w, lab.ptr<int>(i), sizeof(int) * w);
        for (int l = 0; l < n && l < stats_cap; l++) std::memcpy(stats + 5 * l, st.ptr<int>(l), 5 * sizeof(int));
        return n;
    } catch (...) { return -1; }
}
__attribute__((visibility("default"))) int shim_fit_line(const int* xy, int n, int dist, float* line4) {
    try {
        std::vector<cv::Point> pts(n);
        for (int i = 0; i < n; i++) pts[i] = cv::Point(xy[2 * i], xy[2 * i + 1]);
        std::vector<float> line;
        cv::fitLine(pts, line, dist, 0, 0.01, 0.01);
        for (int k = 0; k < 4; k++) line4[k] = line[k];
    } catch (...) { return -1; }
    return 0;
}
__attribute__((visibility("default"))) float shim_fast_atan2(float y, float x) { return cv::fastAtan2(y, x); }
__attribute__((visibility("default"))) int shim_solve2x2(const float* a4, const float* b2, float* x2, double* det) {
    cv::Mat A(2, 2, CV_32FC1), B(2, 1, CV_32FC1), X(2, 1, CV_32FC1);
    A.at<float>(0, 0) = a4[0]; A.at<float>(0, 1) = a4[1]; A.at<float>(1, 0) = a4[2]; A.at<float>(1, 1) = a4[3];
    B.at<float>(0, 0) = b2[0]; B.at<float>(1, 0) = b2[1];
    *det = cv::determinant(A);
    if (*det == 0) return 1;
    cv::solve(A, B, X);
    x2[0] = X.at<float>(0, 0); x2[1] = X.at<float>(1, 0);
    return 0;
}
__attribute__((visibility("default"))) int shim_convert_u8_f32(const unsigned char* src, int n, double alpha, float* dst) {
    cv::Mat s(1, n, CV_8UC1, (void*)src), d;
    s.convertTo(d, CV_32FC1, alpha);
    std::memcpy(dst, d.ptr<float>(0), sizeof(float) * n);
    return 0;
}
__attribute__((visibility("default"))) int shim_bgr2gray(const unsigned char* bgr, int w, int h, unsigned char* gray) {
    cv::Mat s(h, w, CV_8UC3, (void*)bgr), d;
    cv::cvtColor(s, d, cv::COLOR_BGR2GRAY);
    for (int i = 0; i < h; i++) std::memcpy(gray + (size_t)i * w, d.ptr<unsigned char>(i), w);
    return 0;
}
__attribute__((visibility("default"))) int shim_undistort_project(const float* xy, const float* xyz, int n, const float* K9, const float* D5,
                                                                  const double* rvec, const double* tvec, float* und_xy, float* proj_xy) {
    try {
        cv::Mat K(3, 3, CV_32FC1), D(5, 1, CV_32FC1), r(3, 1, CV_64FC1), t(3, 1, CV_64FC1);
        for (int i = 0; i < 9; i++) K.at<float>(i / 3, i % 3) = K9[i];
        for (int i = 0; i < 5; i++) D.at<float>(i, 0) = D5[i];
        for (int i = 0; i < 3; i++) { r.at<double>(i, 0) = rvec[i]; t.at<double>(i, 0) = tvec[i]; }
        std::vector<cv::Point2f> p(n), u, q;
        std::vector<cv::Point3f> o(n);
        for (int i = 0; i < n; i++) { p[i] = cv::Point2f(xy[2 * i], xy[2 * i + 1]); o[i] = cv::Point3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
        cv::undistortPoints(p, u, K, D, cv::noArray(), K);
        cv::projectPoints(o, r, t, K, D, q);
        for (int i = 0; i < n; i++) { und_xy[2 * i] = u[i].x; und_xy[2 * i + 1] = u[i].y; proj_xy[2 * i] = q[i].x; proj_xy[2 * i + 1] = q[i].y; }
    } catch (...) { return -1; }
    return 0;
}

}  // extern "C"
