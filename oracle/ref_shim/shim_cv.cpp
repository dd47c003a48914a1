// TEST INFRASTRUCTURE ONLY -- never linked into the product (cylindertag_b200/).
//
// Implementations behind oracle/ref_shim/opencv2/core.hpp: restatements of the published OpenCV 4.x algorithms for the
// entry points the reference's detect / estimatePose path calls, each for exactly the
// argument types the reference passes.  OpenCV is a third-party dependency that is absent from /root/reference and
// has no C++ SDK in this container (README.md:20 of the reference: tested on 4.5.3; build/CMakeCache.txt:468: 4.6.0;
// this container's Python wheel: 4.13.0, which is what the restatements are pinned to -- tests/test_ref_shim.py runs
// every routine against cv2 and the whole reference with the routines routed to cv2 through shim_backend).
#include <opencv2/core.hpp>

#include <cstdio>
#include <fstream>
#include <sstream>

namespace {
shim_backend g_backend = {};  // zero = built-in restatements
bool g_have_backend = false;
}  // namespace

extern "C" void shim_set_backend(const shim_backend* b) {
    if (b) { g_backend = *b; g_have_backend = true; }
    else { g_backend = shim_backend(); g_have_backend = false; }
}

namespace cv {

// ---------------------------------------------------------------------------------------------------------------------
// Mat::convertTo.  u8 -> f32 (the reference's only use, CylinderTag.cpp:80,101) follows cvt_32f: float(v) * float(alpha)
// + float(beta) evaluated in binary32 (the fp64-then-round form differs on 122 k pixels of test.bmp).
void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const {
    int ddepth = CV_MAT_DEPTH(rtype);
    Mat out(rows, cols, CV_MAKETYPE(ddepth, channels()));
    const int n = cols * channels();
    if (depth() == CV_8U && ddepth == CV_32F) {
        const float a = (float)alpha, b = (float)beta;
        for (int i = 0; i < rows; i++) {
            const uchar* s = ptr<uchar>(i);
            float* d = out.ptr<float>(i);
            if (g_backend.convert_u8_f32 && beta == 0) {
                if (g_backend.convert_u8_f32(s, n, alpha, d) != 0) throw Exception("convertTo backend failed");
                continue;
            }
            for (int j = 0; j < n; j++) d[j] = (float)s[j] * a + b;
        }
    } else if (depth() == ddepth && alpha == 1 && beta == 0) {
        out = clone();
    } else if (depth() == CV_32F && ddepth == CV_64F) {
        for (int i = 0; i < rows; i++) for (int j = 0; j < n; j++) out.ptr<double>(i)[j] = ptr<float>(i)[j] * alpha + beta;
    } else if (depth() == CV_64F && ddepth == CV_32F) {
        for (int i = 0; i < rows; i++) for (int j = 0; j < n; j++) out.ptr<float>(i)[j] = (float)(ptr<double>(i)[j] * alpha + beta);
    } else {
        throw Exception("shim Mat::convertTo: conversion not used by the reference");
    }
    dst = out;
}

// cvtColor: GRAY2RGB (CylinderTag.cpp:70,214; its result is only drawn on) and BGR2GRAY (main.cpp:36,54;
// (3735 B + 19235 G + 9798 R + 2^14) >> 15).
void cvtColor(const Mat& src, Mat& dst, int code) {
    if (code == COLOR_GRAY2RGB) {
        if (src.type() != CV_8UC1) throw Exception("shim cvtColor(GRAY2RGB): 8UC1 input expected");
        Mat out(src.rows, src.cols, CV_8UC3);
        for (int i = 0; i < src.rows; i++) {
            const uchar* s = src.ptr<uchar>(i);
            uchar* d = out.ptr<uchar>(i);
            for (int j = 0; j < src.cols; j++) d[3 * j] = d[3 * j + 1] = d[3 * j + 2] = s[j];
        }
        dst = out;
    } else if (code == COLOR_BGR2GRAY) {
        if (src.type() != CV_8UC3) throw Exception("shim cvtColor(BGR2GRAY): 8UC3 input expected");
        Mat out(src.rows, src.cols, CV_8UC1);
        for (int i = 0; i < src.rows; i++) {
            const uchar* s = src.ptr<uchar>(i);
            uchar* d = out.ptr<uchar>(i);
            for (int j = 0; j < src.cols; j++)
                d[j] = (uchar)((3735 * s[3 * j] + 19235 * s[3 * j + 1] + 9798 * s[3 * j + 2] + (1 << 14)) >> 15);
        }
        dst = out;
    } else {
        throw Exception("shim cvtColor: code not used by the reference");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// resize, INTER_CUBIC, 8UC1 (CylinderTag.cpp:79).  Restates imgproc/resize.cpp's generic path for that case:
// per destination column/row the source position ((d + 0.5) * scale - 0.5) in float, Keys' bicubic weights with
// A = -0.75 in float, rounded to 11-bit fixed point (cvRound), replicate border; horizontal pass in exact int32;
// vertical pass (VResizeCubicVec_32s8u): S0*b0 + (S1*b1 + (S2*b2 + S3*b3)) in binary32 with b = beta * 2^-22, rounded
// half to even.  cv2 4.13 gives this float form on every column, also where the width is not a multiple of the vector
// length (the integer form (v + 2^21) >> 22 differs on ties; measured, tests/test_ref_pinning.py).
// Non-integer scales (odd source sizes): this restatement equals cv2 with cv::setUseOptimized(false); with CPU dispatch
// on, the library's own result differs by 1 DN on ~5 % of the pixels, i.e. OpenCV itself has no single answer there,
// which is why every config and the product's C ABI keep to even sizes (exact 2x, identical either way).
static inline void cubic_coeffs(float x, float* c) {
    const float A = -0.75f;
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
}

static void resize_cubic_u8_native(const uchar* src, int sw, int sh, size_t sstep, uchar* dst, int dw, int dh, size_t dstep) {
    const double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> alpha((size_t)dw * 4), beta((size_t)dh * 4);
    float c[4];
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cvFloor(fx);
        fx -= sx;
        xofs[dx] = sx;
        cubic_coeffs(fx, c);
        for (int k = 0; k < 4; k++) alpha[(size_t)dx * 4 + k] = saturate_cast<short>(c[k] * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cvFloor(fy);
        fy -= sy;
        yofs[dy] = sy;
        cubic_coeffs(fy, c);
        for (int k = 0; k < 4; k++) beta[(size_t)dy * 4 + k] = saturate_cast<short>(c[k] * 2048.f);
    }
    // horizontal pass of every source row once (exact integers, so caching order is irrelevant)
    std::vector<int> hbuf((size_t)sh * dw);
    for (int y = 0; y < sh; y++) {
        const uchar* S = src + (size_t)y * sstep;
        int* D = &hbuf[(size_t)y * dw];
        for (int dx = 0; dx < dw; dx++) {
            const short* a = &alpha[(size_t)dx * 4];
            int v = 0;
            for (int j = 0; j < 4; j++) {
                int sxj = xofs[dx] - 1 + j;
                sxj = sxj < 0 ? 0 : sxj >= sw ? sw - 1 : sxj;
                v += S[sxj] * a[j];
            }
            D[dx] = v;
        }
    }
    for (int dy = 0; dy < dh; dy++) {
        const int* S[4];
        for (int k = 0; k < 4; k++) {
            int sy = yofs[dy] - 1 + k;
            sy = sy < 0 ? 0 : sy >= sh ? sh - 1 : sy;
            S[k] = &hbuf[(size_t)sy * dw];
        }
        const short* b = &beta[(size_t)dy * 4];
        const float scale = 1.f / (2048 * 2048);
        const float b0 = b[0] * scale, b1 = b[1] * scale, b2 = b[2] * scale, b3 = b[3] * scale;
        uchar* D = dst + (size_t)dy * dstep;
        for (int x = 0; x < dw; x++) {
            float t = (float)S[3][x] * b3;  // built with -ffp-contract=off: every step rounds to binary32
            t = (float)S[2][x] * b2 + t;
            t = (float)S[1][x] * b1 + t;
            t = (float)S[0][x] * b0 + t;
            int r = cvRound(t);
            D[x] = (uchar)(r < 0 ? 0 : r > 255 ? 255 : r);
        }
    }
}

void resize(const Mat& src, Mat& dst, Size dsize, double fx, double fy, int interpolation) {
    if (interpolation != INTER_CUBIC || src.type() != CV_8UC1) throw Exception("shim resize: only INTER_CUBIC on 8UC1 (CylinderTag.cpp:79)");
    if (dsize.empty()) dsize = Size(saturate_cast<int>(src.cols * fx), saturate_cast<int>(src.rows * fy));
    if (dsize.empty()) throw Exception("shim resize: empty destination");
    Mat out(dsize.height, dsize.width, CV_8UC1);
    if (g_backend.resize_cubic_u8) {
        if (g_backend.resize_cubic_u8(src.data, src.cols, src.rows, src.step, out.data, out.cols, out.rows, out.step) != 0)
            throw Exception("resize backend failed");
    } else {
        resize_cubic_u8_native(src.data, src.cols, src.rows, src.step, out.data, out.cols, out.rows, out.step);
    }
    dst = out;
}

// minMaxLoc on a 32F (or 8U) single-channel view (corner_detector.cpp:46,58,60).  Values only feed (float) casts.
void minMaxLoc(const Mat& src, double* minVal, double* maxVal, Point* minLoc, Point* maxLoc) {
    if (src.empty() || src.channels() != 1) throw Exception("shim minMaxLoc: single-channel non-empty input expected");
    double mn = DBL_MAX, mx = -DBL_MAX;
    Point pmn, pmx;
    for (int i = 0; i < src.rows; i++)
        for (int j = 0; j < src.cols; j++) {
            double v;
            switch (src.depth()) {
                case CV_32F: v = src.at<float>(i, j); break;
                case CV_8U: v = src.at<uchar>(i, j); break;
                case CV_32S: v = src.at<int>(i, j); break;
                case CV_64F: v = src.at<double>(i, j); break;
                default: throw Exception("shim minMaxLoc: depth");
            }
            if (v < mn) { mn = v; pmn = Point(j, i); }
            if (v > mx) { mx = v; pmx = Point(j, i); }
        }
    if (minVal) *minVal = mn;
    if (maxVal) *maxVal = mx;
    if (minLoc) *minLoc = pmn;
    if (maxLoc) *maxLoc = pmx;
}

// ---------------------------------------------------------------------------------------------------------------------
// connectedComponentsWithStats(8-connectivity, CV_32S, CCL_BBDT) (corner_detector.cpp:82).  The partition into
// 8-connected components is unique; what the algorithm choice fixes is the label ORDER, and the reference's greedy
// pairing (corner_detector.cpp:482-557) depends on it.  BBDT scans 2x2 blocks in raster order, gives a block with no
// labelled neighbour the next provisional label, merges towards the smaller label and flattens in increasing order, so
// final labels ascend with the raster index of each component's first 2x2 block (unaffected by OpenCV's
// row-striped parallel variant, whose stripes are ordered the same way).
static int ccl_native(const uchar* img, int w, int h, size_t step, int* labels) {
    std::vector<int> parent(1, 0);
    auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    for (int y = 0; y < h; y++) {
        const uchar* row = img + (size_t)y * step;
        int* L = labels + (size_t)y * w;
        const int* U = y ? L - w : nullptr;
        for (int x = 0; x < w; x++) {
            if (!row[x]) { L[x] = 0; continue; }
            int cand[4], nc = 0;
            if (x && L[x - 1]) cand[nc++] = L[x - 1];
            if (U) {
                if (x && U[x - 1]) cand[nc++] = U[x - 1];
                if (U[x]) cand[nc++] = U[x];
                if (x + 1 < w && U[x + 1]) cand[nc++] = U[x + 1];
            }
            if (!nc) { parent.push_back((int)parent.size()); L[x] = (int)parent.size() - 1; continue; }
            int r = find(cand[0]);
            for (int k = 1; k < nc; k++) {
                int q = find(cand[k]);
                if (q != r) { if (q < r) std::swap(q, r); parent[q] = r; }
            }
            L[x] = r;
        }
    }
    const int bw = (w + 1) / 2;
    const int np = (int)parent.size();
    std::vector<long long> key(np, LLONG_MAX);
    for (int y = 0; y < h; y++) {
        int* L = labels + (size_t)y * w;
        for (int x = 0; x < w; x++)
            if (L[x]) {
                int r = find(L[x]);
                L[x] = r;
                long long k = (long long)(y >> 1) * bw + (x >> 1);
                if (k < key[r]) key[r] = k;
            }
    }
    std::vector<int> roots;
    for (int p = 1; p < np; p++) if (parent[p] == p && key[p] != LLONG_MAX) roots.push_back(p);
    std::sort(roots.begin(), roots.end(), [&](int a, int b) { return key[a] < key[b]; });
    std::vector<int> final_label(np, 0);
    for (size_t i = 0; i < roots.size(); i++) final_label[roots[i]] = (int)i + 1;
    for (size_t i = 0; i < (size_t)w * h; i++) labels[i] = final_label[labels[i]];
    return (int)roots.size() + 1;
}

int connectedComponentsWithStats(const Mat& image, Mat& labels, Mat& stats, Mat& centroids, int connectivity, int ltype, int ccltype) {
    if (image.type() != CV_8UC1 || connectivity != 8 || ltype != CV_32S || (ccltype != CCL_BBDT && ccltype != CCL_DEFAULT && ccltype != CCL_GRANA))
        throw Exception("shim connectedComponentsWithStats: only (8UC1, 8, CV_32S, CCL_BBDT) (corner_detector.cpp:82)");
    const int w = image.cols, h = image.rows;
    Mat lab(h, w, CV_32SC1);
    int n;
    if (g_backend.ccl_bbdt) n = g_backend.ccl_bbdt(image.data, w, h, image.step, lab.ptr<int>());
    else n = ccl_native(image.data, w, h, image.step, lab.ptr<int>());
    if (n <= 0) throw Exception("ccl backend failed");
    Mat st(n, 5, CV_32SC1), cen(n, 2, CV_64FC1);
    std::vector<long long> sx(n, 0), sy(n, 0);
    for (int l = 0; l < n; l++) { st.at<int>(l, CC_STAT_LEFT) = INT_MAX; st.at<int>(l, CC_STAT_TOP) = INT_MAX; st.at<int>(l, 2) = INT_MIN; st.at<int>(l, 3) = INT_MIN; }
    for (int y = 0; y < h; y++) {
        const int* L = lab.ptr<int>(y);
        for (int x = 0; x < w; x++) {
            int l = L[x];
            int* s = st.ptr<int>(l);
            if (x < s[0]) s[0] = x;
            if (y < s[1]) s[1] = y;
            if (x > s[2]) s[2] = x;
            if (y > s[3]) s[3] = y;
            s[4]++;
            sx[l] += x; sy[l] += y;
        }
    }
    for (int l = 0; l < n; l++) {
        int* s = st.ptr<int>(l);
        if (s[4] == 0) { s[0] = s[1] = s[2] = s[3] = 0; cen.at<double>(l, 0) = cen.at<double>(l, 1) = 0; continue; }
        s[2] = s[2] - s[0] + 1;
        s[3] = s[3] - s[1] + 1;
        cen.at<double>(l, 0) = (double)sx[l] / s[4];
        cen.at<double>(l, 1) = (double)sy[l] / s[4];
    }
    labels = lab; stats = st; centroids = cen;
    return n;
}

// ---------------------------------------------------------------------------------------------------------------------
// fitLine on 2-D points (imgproc/linefit.cpp), DIST_L2 (corner_detector.cpp:136,151,163) and DIST_WELSCH (:358).
// Points are converted to float first.  L2: fp64 sums of float products, t = float(atan2(2 dxy, dx2 - dy2)) / 2,
// line = (cosf t, sinf t, mean x, mean y).  Robust distances: a fresh MWC generator RNG(-1) per call, 20 restarts from
// min(n, 10) distinct random points, up to 30 reweighting passes each; the best-so-far test sits inside the pass loop
// (4.13) as well as after it.
static void fit_line_wods(const Point2f* p, int n, const float* w, float* line) {
    double x = 0, y = 0, x2 = 0, y2 = 0, xy = 0, sw = 0;
    if (!w) {
        for (int i = 0; i < n; i++) {
            x += p[i].x; y += p[i].y;
            x2 += p[i].x * p[i].x; y2 += p[i].y * p[i].y; xy += p[i].x * p[i].y;
        }
        sw = (float)n;
    } else {
        for (int i = 0; i < n; i++) {
            x += w[i] * p[i].x; y += w[i] * p[i].y;
            x2 += w[i] * p[i].x * p[i].x; y2 += w[i] * p[i].y * p[i].y; xy += w[i] * p[i].x * p[i].y;
            sw += w[i];
        }
    }
    x /= sw; y /= sw; x2 /= sw; y2 /= sw; xy /= sw;
    double dx2 = x2 - x * x, dy2 = y2 - y * y, dxy = xy - x * y;
    float t = (float)std::atan2(2 * dxy, dx2 - dy2) / 2;
    line[0] = (float)std::cos(t);  // float overloads: cosf / sinf
    line[1] = (float)std::sin(t);
    line[2] = (float)x;
    line[3] = (float)y;
}

static double calc_dist2d(const Point2f* p, int n, const float* line, float* dist) {
    float px = line[2], py = line[3], nx = line[1], ny = -line[0];
    double sum = 0;
    for (int j = 0; j < n; j++) {
        float x = p[j].x - px, y = p[j].y - py;
        dist[j] = (float)std::fabs(nx * x + ny * y);
        sum += dist[j];
    }
    return sum;
}

struct MwcRng {
    uint64_t state;
    explicit MwcRng(uint64_t s) : state(s) {}
    unsigned next() { state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32); return (unsigned)state; }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

static void fit_line_2d(const Point2f* p, int n, int dist, float param, float reps, float aeps, float* line) {
    std::memset(line, 0, 4 * sizeof(float));
    if (dist == DIST_L2) { fit_line_wods(p, n, nullptr, line); return; }
    if (dist != DIST_WELSCH) throw Exception("shim fitLine: only DIST_L2 and DIST_WELSCH are used by the reference");
    const double EPS = n * FLT_EPSILON;
    const float rdelta = reps != 0 ? reps : 1.0f, adelta = aeps != 0 ? aeps : 0.01f;
    double min_err = DBL_MAX, err = 0;
    MwcRng rng((uint64_t)-1);
    std::vector<float> wr((size_t)n * 2);
    float* w = wr.data();
    float* r = w + n;
    float cur[4], prev[4] = {0, 0, 0, 0};
    const float c = param == 0 ? 1 / 2.9846f : 1 / param;
    for (int k = 0; k < 20; k++) {
        int first = 1;
        for (int i = 0; i < n; i++) w[i] = 0.f;
        for (int i = 0; i < std::min(n, 10);) {
            int j = rng.uniform(0, n);
            if (w[j] < FLT_EPSILON) { w[j] = 1.f; i++; }
        }
        fit_line_wods(p, n, w, cur);
        for (int i = 0; i < 30; i++) {
            double sum_w = 0;
            if (first) first = 0;
            else {
                double t = cur[0] * prev[0] + cur[1] * prev[1];
                t = std::max(t, -1.);
                t = std::min(t, 1.);
                if (std::fabs(std::acos(t)) < adelta) {
                    float x = (float)std::fabs(cur[2] - prev[2]), y = (float)std::fabs(cur[3] - prev[3]);
                    float d = x > y ? x : y;
                    if (d < rdelta) break;
                }
            }
            err = calc_dist2d(p, n, cur, r);
            if (err < min_err) {
                min_err = err;
                std::memcpy(line, cur, sizeof(cur));
                if (err < EPS) break;
            }
            for (int j = 0; j < n; j++) w[j] = (float)std::exp(-r[j] * r[j] * c * c);  // float overload: expf
            for (int j = 0; j < n; j++) sum_w += w[j];
            if (std::fabs(sum_w) > FLT_EPSILON) {
                sum_w = 1. / sum_w;
                for (int j = 0; j < n; j++) w[j] = (float)(w[j] * sum_w);
            } else {
                for (int j = 0; j < n; j++) w[j] = 1.f;
            }
            std::memcpy(prev, cur, sizeof(cur));
            fit_line_wods(p, n, w, cur);
        }
        if (err < min_err) {
            min_err = err;
            std::memcpy(line, cur, sizeof(cur));
            if (err < EPS) break;
        }
    }
}

void fitLine(const std::vector<Point2f>& points, std::vector<float>& line, int distType, double param, double reps, double aeps) {
    if (points.empty()) throw Exception("shim fitLine: empty input");
    float out[4];
    if (g_backend.fit_line) {
        if (g_backend.fit_line(&points[0].x, (int)points.size(), distType, param, reps, aeps, out) != 0) throw Exception("fitLine backend failed");
    } else {
        fit_line_2d(points.data(), (int)points.size(), distType, (float)param, (float)reps, (float)aeps, out);
    }
    line.assign(out, out + 4);
}

void fitLine(const std::vector<Point>& points, std::vector<float>& line, int distType, double param, double reps, double aeps) {
    std::vector<Point2f> pf(points.size());
    for (size_t i = 0; i < points.size(); i++) pf[i] = Point2f((float)points[i].x, (float)points[i].y);
    fitLine(pf, line, distType, param, reps, aeps);
}

Scalar sum(const Mat& src) {
    Scalar s;
    const int cn = src.channels();
    if (cn > 4) throw Exception("shim sum: channels");
    for (int i = 0; i < src.rows; i++)
        for (int j = 0; j < src.cols * cn; j++) {
            double v;
            switch (src.depth()) {
                case CV_8U: v = src.ptr<uchar>(i)[j]; break;
                case CV_32S: v = src.ptr<int>(i)[j]; break;
                case CV_32F: v = src.ptr<float>(i)[j]; break;
                case CV_64F: v = src.ptr<double>(i)[j]; break;
                default: throw Exception("shim sum: depth");
            }
            s[j % cn] += v;
        }
    return s;
}

// determinant / solve for 2x2 CV_32F (core/lapack.cpp, the closed forms).
double determinant(const Mat& m) {
    if (m.type() != CV_32FC1 || m.rows != 2 || m.cols != 2) throw Exception("shim determinant: 2x2 CV_32F only");
    return (double)m.at<float>(0, 0) * m.at<float>(1, 1) - (double)m.at<float>(0, 1) * m.at<float>(1, 0);
}

bool solve(const Mat& A, const Mat& B, Mat& X, int flags) {
    if (flags != DECOMP_LU || A.type() != CV_32FC1 || A.rows != 2 || A.cols != 2 || B.type() != CV_32FC1 || B.rows != 2 || B.cols != 1)
        throw Exception("shim solve: 2x2 CV_32F, DECOMP_LU only");
    if (X.type() != CV_32FC1 || X.rows != 2 || X.cols != 1 || X.data == A.data || X.data == B.data) X = Mat(2, 1, CV_32FC1);
    double d = determinant(A);
    if (d == 0) return false;
    d = 1. / d;
    const float b0 = B.at<float>(0, 0), b1 = B.at<float>(1, 0);
    float t = (float)(((double)b0 * A.at<float>(1, 1) - (double)b1 * A.at<float>(0, 1)) * d);
    X.at<float>(1, 0) = (float)(((double)b1 * A.at<float>(0, 0) - (double)b0 * A.at<float>(1, 0)) * d);
    X.at<float>(0, 0) = t;
    return true;
}

// fastAtan2 (core/mathfuncs_core): degrees in [0, 360), 7th-order odd polynomial on the folded octant.  The library's
// vector body contracts with FMA; this scalar form agrees to ~3e-5 degrees.  It only feeds the 45 / 135
// degree ordering decision (corner_detector.cpp:1028-1034).
float fastAtan2(float y, float x) {
    if (g_backend.fast_atan2) return g_backend.fast_atan2(y, x);
    const float p1 = 0.9997878412794807f * (float)(180 / CV_PI), p3 = -0.3258083974640975f * (float)(180 / CV_PI),
                p5 = 0.1555786518463281f * (float)(180 / CV_PI), p7 = -0.04432655554792128f * (float)(180 / CV_PI);
    float ax = std::abs(x), ay = std::abs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------------------------------------------------
// calib3d.  projectPoints / undistortPoints restate the published pinhole + Brown-Conrady (k1 k2 p1 p2 k3) model;
// solvePnP(EPNP) is only available through the cv2 backend (the EPnP control-point solver is not restated here).
static void cam_params(const Mat& K, const Mat& D, float* k9, std::vector<float>& d) {
    if (K.rows != 3 || K.cols != 3) throw Exception("shim calib3d: cameraMatrix must be 3x3");
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) k9[3 * i + j] = K.depth() == CV_32F ? K.at<float>(i, j) : (float)K.at<double>(i, j);
    d.clear();
    if (!D.empty()) {
        int n = D.rows * D.cols;
        for (int i = 0; i < n; i++) {
            int r = D.cols == 1 ? i : 0, c = D.cols == 1 ? 0 : i;
            d.push_back(D.depth() == CV_32F ? D.at<float>(r, c) : (float)D.at<double>(r, c));
        }
    }
}

static void rodrigues(const double* r, double* R) {
    double th = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < DBL_EPSILON) { for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0); return; }
    double x = r[0] / th, y = r[1] / th, z = r[2] / th, c = std::cos(th), s = std::sin(th), c1 = 1 - c;
    R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

bool solvePnP(const std::vector<Point3f>& obj, const std::vector<Point2f>& img, const Mat& K, const Mat& D, Mat& rvec, Mat& tvec,
              bool useExtrinsicGuess, int flags) {
    if (flags != SOLVEPNP_EPNP || useExtrinsicGuess) throw Exception("shim solvePnP: SOLVEPNP_EPNP without a guess only (pose_estimation.cpp:96)");
    if (!g_backend.solve_pnp_epnp) throw Exception("shim solvePnP: needs the cv2 backend (shim_set_backend)");
    if (obj.size() != img.size() || obj.size() < 4) throw Exception("shim solvePnP: point counts");
    float k9[9];
    std::vector<float> d;
    cam_params(K, D, k9, d);
    Mat r(3, 1, CV_64FC1), t(3, 1, CV_64FC1);
    if (g_backend.solve_pnp_epnp(&obj[0].x, &img[0].x, (int)obj.size(), k9, d.data(), (int)d.size(), r.ptr<double>(), t.ptr<double>()) != 0)
        return false;
    rvec = r; tvec = t;
    return true;
}

void undistortPoints(const std::vector<Point2f>& src, std::vector<Point2f>& dst, const Mat& K, const Mat& D, NoArray, const Mat& P) {
    float k9[9], p9[9];
    std::vector<float> d, unused;
    cam_params(K, D, k9, d);
    cam_params(P, Mat(), p9, unused);
    std::vector<Point2f> out(src.size());
    if (src.empty()) { dst = out; return; }
    if (g_backend.undistort_points) {
        for (int i = 0; i < 9; i++) if (k9[i] != p9[i]) throw Exception("shim undistortPoints backend: P must equal the camera matrix");
        if (g_backend.undistort_points(&src[0].x, (int)src.size(), k9, d.data(), (int)d.size(), &out[0].x) != 0) throw Exception("undistortPoints backend failed");
        dst = out;
        return;
    }
    d.resize(5, 0.f);
    const double fx = k9[0], fy = k9[4], cx = k9[2], cy = k9[5], k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
    for (size_t i = 0; i < src.size(); i++) {
        double x0 = (src[i].x - cx) / fx, y0 = (src[i].y - cy) / fy, x = x0, y = y0;
        for (int it = 0; it < 5; it++) {  // cvUndistortPointsInternal's default criteria: 5 fixed-point passes
            double r2 = x * x + y * y, icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
            if (icdist < 0) { x = x0; y = y0; break; }
            double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x), dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
            x = (x0 - dx) * icdist;
            y = (y0 - dy) * icdist;
        }
        out[i] = Point2f((float)(x * p9[0] + p9[2]), (float)(y * p9[4] + p9[5]));
    }
    dst = out;
}

void projectPoints(const std::vector<Point3f>& obj, const Mat& rvec, const Mat& tvec, const Mat& K, const Mat& D, std::vector<Point2f>& img) {
    float k9[9];
    std::vector<float> d;
    cam_params(K, D, k9, d);
    double r[3], t[3];
    for (int i = 0; i < 3; i++) {
        r[i] = rvec.depth() == CV_64F ? rvec.at<double>(i) : rvec.at<float>(i);
        t[i] = tvec.depth() == CV_64F ? tvec.at<double>(i) : tvec.at<float>(i);
    }
    std::vector<Point2f> out(obj.size());
    if (obj.empty()) { img = out; return; }
    if (g_backend.project_points) {
        if (g_backend.project_points(&obj[0].x, (int)obj.size(), r, t, k9, d.data(), (int)d.size(), &out[0].x) != 0) throw Exception("projectPoints backend failed");
        img = out;
        return;
    }
    d.resize(5, 0.f);
    double R[9];
    rodrigues(r, R);
    const double fx = k9[0], fy = k9[4], cx = k9[2], cy = k9[5], k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
    for (size_t i = 0; i < obj.size(); i++) {
        double X = R[0] * obj[i].x + R[1] * obj[i].y + R[2] * obj[i].z + t[0];
        double Y = R[3] * obj[i].x + R[4] * obj[i].y + R[5] * obj[i].z + t[1];
        double Z = R[6] * obj[i].x + R[7] * obj[i].y + R[8] * obj[i].z + t[2];
        double x = X / Z, y = Y / Z, r2 = x * x + y * y, cd = 1 + ((k3 * r2 + k2) * r2 + k1) * r2;
        double xd = x * cd + 2 * p1 * x * y + p2 * (r2 + 2 * x * x), yd = y * cd + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
        out[i] = Point2f((float)(xd * fx + cx), (float)(yd * fy + cy));
    }
    img = out;
}

// ---------------------------------------------------------------------------------------------------------------------
// FileStorage: the "%YAML:1.0" subset cameraParams.yml uses -- top-level keys holding !!opencv-matrix maps with
// rows, cols, dt and a flow-sequence data list that may span lines.
FileStorage::FileStorage(const std::string& path, int mode) : opened(false) {
    if (mode != READ) return;
    std::ifstream f(path);
    if (!f.is_open()) return;
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string txt = ss.str();
    opened = true;
    size_t pos = 0;
    while ((pos = txt.find("!!opencv-matrix", pos)) != std::string::npos) {
        size_t ls = txt.rfind('\n', pos);
        ls = ls == std::string::npos ? 0 : ls + 1;
        size_t colon = txt.find(':', ls);
        std::string key = txt.substr(ls, colon - ls);
        key.erase(0, key.find_first_not_of(" \t"));
        key.erase(key.find_last_not_of(" \t") + 1);
        FileNode n;
        n.ok = true;
        size_t next = txt.find("!!opencv-matrix", pos + 1);
        const std::string body = txt.substr(pos, next == std::string::npos ? std::string::npos : next - pos);
        auto field = [&](const char* name) -> std::string {
            size_t p = body.find(name);
            if (p == std::string::npos) return "";
            p = body.find(':', p) + 1;
            size_t e = body.find('\n', p);
            std::string v = body.substr(p, e - p);
            v.erase(0, v.find_first_not_of(" \t\""));
            v.erase(v.find_last_not_of(" \t\"\r") + 1);
            return v;
        };
        n.rows = std::atoi(field("rows").c_str());
        n.cols = std::atoi(field("cols").c_str());
        std::string dt = field("dt");
        n.dt = dt.empty() ? 'f' : dt[0];
        size_t b = body.find('[', body.find("data")), e = body.find(']', b);
        std::string list = body.substr(b + 1, e - b - 1);
        for (char& ch : list) if (ch == ',') ch = ' ';
        std::stringstream ls2(list);
        double v;
        while (ls2 >> v) n.values.push_back(v);
        nodes.push_back(std::make_pair(key, n));
        pos += 15;
    }
}

FileNode FileStorage::operator[](const std::string& key) const {
    for (const auto& kv : nodes) if (kv.first == key) return kv.second;
    return FileNode();
}

void operator>>(const FileNode& n, Mat& m) {
    if (!n.ok) { m = Mat(); return; }
    if ((int)n.values.size() != n.rows * n.cols) throw Exception("shim FileStorage: matrix data count");
    int type = n.dt == 'd' ? CV_64FC1 : n.dt == 'i' ? CV_32SC1 : n.dt == 'u' ? CV_8UC1 : CV_32FC1;
    Mat out(n.rows, n.cols, type);
    for (int i = 0; i < n.rows; i++)
        for (int j = 0; j < n.cols; j++) {
            double v = n.values[(size_t)i * n.cols + j];
            switch (type) {
                case CV_64FC1: out.at<double>(i, j) = v; break;
                case CV_32SC1: out.at<int>(i, j) = (int)v; break;
                case CV_8UC1: out.at<uchar>(i, j) = (uchar)v; break;
                default: out.at<float>(i, j) = (float)v;
            }
        }
    m = out;
}

void imshow(const std::string&, const Mat&) {}

}  // namespace cv
