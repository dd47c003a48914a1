// TEST INFRASTRUCTURE ONLY -- never linked into the product (cylindertag_b200/).
//
// A minimal stand-in for the parts of OpenCV 4 that the reference's detect / estimatePose path calls, so that the
// UNMODIFIED reference sources (/root/reference/corner_detector.cpp, CylinderTag.cpp, pose_estimation.cpp) compile and
// run in a container that has no OpenCV C++ SDK (oracle/build_ref.py -> oracle/_ref/libctag_ref.so).  It is NOT
// OpenCV: only the entry points listed in SURVEY.md 8(c) exist, and each numeric routine restates the published
// OpenCV 4.x algorithm for exactly the argument types the reference passes.  Every routine can also be routed to the
// real library through cv2 callbacks (shim_backend, below) -- tests/test_ref_shim.py runs both ways and requires
// identical results, which is what pins the restatements.
//
// Call sites served (reference file:line): resize CylinderTag.cpp:79; Mat::convertTo :80,101; cvtColor :70,214;
// minMaxLoc corner_detector.cpp:46,58,60; connectedComponentsWithStats :82; fitLine :136,151,163,358; sum :246;
// norm :285,288,337; determinant/solve :370-371,1111-1151; fastAtan2 :1028; solvePnP pose_estimation.cpp:96;
// undistortPoints :109; projectPoints CylinderTag.cpp:234; FileStorage :193; drawing/highgui :236-245 (no-ops).
//
// Decision C-1: Mat(rows, cols, type) zero-fills its buffer, so the threshold's border tile
// ring, which the reference reads uninitialised (corner_detector.cpp:31-34,71), is 0.
#pragma once
#ifndef CTAG_REF_SHIM_OPENCV_CORE_HPP
#define CTAG_REF_SHIM_OPENCV_CORE_HPP

#include <algorithm>
#include <array>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795

namespace cv {

typedef unsigned char uchar;

class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string& m) : std::runtime_error(m) {}
};

// cvRound: round half to even (SSE cvtsd2si), the rounding every saturate_cast to an integer type uses.
inline int cvRound(double v) { return (int)std::lrint(v); }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }

template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <> inline short saturate_cast<short>(double v) {
    int i = cvRound(v);
    return (short)(i < SHRT_MIN ? SHRT_MIN : i > SHRT_MAX ? SHRT_MAX : i);
}
template <> inline uchar saturate_cast<uchar>(double v) {
    int i = cvRound(v);
    return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i);
}
template <typename T, typename U> struct PtCast { static T cast(U v) { return saturate_cast<T>((double)v); } };
template <typename T> struct PtCast<T, T> { static T cast(T v) { return v; } };
template <> struct PtCast<float, int> { static float cast(int v) { return (float)v; } };
template <> struct PtCast<double, int> { static double cast(int v) { return (double)v; } };
template <> struct PtCast<double, float> { static double cast(float v) { return (double)v; } };
template <> struct PtCast<float, double> { static float cast(double v) { return (float)v; } };

template <typename T> class Point_ {
public:
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> operator Point_<U>() const { return Point_<U>(PtCast<U, T>::cast(x), PtCast<U, T>::cast(y)); }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> inline Point_<T> operator-(const Point_<T>& a) { return Point_<T>(-a.x, -a.y); }
template <typename T> inline Point_<T>& operator+=(Point_<T>& a, const Point_<T>& b) { a.x += b.x; a.y += b.y; return a; }
template <typename T> inline Point_<T>& operator-=(Point_<T>& a, const Point_<T>& b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename T> inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> inline bool operator!=(const Point_<T>& a, const Point_<T>& b) { return !(a == b); }
// OpenCV evaluates point * scalar in the promoted type of (T, scalar) and casts back with saturate_cast<T>.
#define CTAG_SHIM_PT_SCALAR(S)                                                                                         \
    template <typename T> inline Point_<T> operator*(const Point_<T>& a, S b) {                                        \
        return Point_<T>(PtCast<T, decltype(a.x * b)>::cast(a.x * b), PtCast<T, decltype(a.x * b)>::cast(a.y * b));    \
    }                                                                                                                  \
    template <typename T> inline Point_<T> operator*(S b, const Point_<T>& a) { return a * b; }                        \
    template <typename T> inline Point_<T> operator/(const Point_<T>& a, S b) {                                        \
        return Point_<T>(PtCast<T, decltype(a.x / b)>::cast(a.x / b), PtCast<T, decltype(a.x / b)>::cast(a.y / b));    \
    }
CTAG_SHIM_PT_SCALAR(int)
CTAG_SHIM_PT_SCALAR(float)
CTAG_SHIM_PT_SCALAR(double)
#undef CTAG_SHIM_PT_SCALAR

template <typename T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }

template <typename T> class Point3_ {
public:
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
    template <typename U> operator Point3_<U>() const {
        return Point3_<U>(PtCast<U, T>::cast(x), PtCast<U, T>::cast(y), PtCast<U, T>::cast(z));
    }
};
typedef Point3_<int> Point3i;
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
template <typename T> inline Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> inline Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
#define CTAG_SHIM_PT3_SCALAR(S)                                                                                        \
    template <typename T> inline Point3_<T> operator*(const Point3_<T>& a, S b) {                                      \
        typedef decltype(a.x * b) W;                                                                                   \
        return Point3_<T>(PtCast<T, W>::cast(a.x * b), PtCast<T, W>::cast(a.y * b), PtCast<T, W>::cast(a.z * b));      \
    }                                                                                                                  \
    template <typename T> inline Point3_<T> operator*(S b, const Point3_<T>& a) { return a * b; }
CTAG_SHIM_PT3_SCALAR(int)
CTAG_SHIM_PT3_SCALAR(float)
CTAG_SHIM_PT3_SCALAR(double)
#undef CTAG_SHIM_PT3_SCALAR

template <typename T> class Size_ {
public:
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool empty() const { return width <= 0 || height <= 0; }
};
typedef Size_<int> Size;

template <typename T> class Rect_ {
public:
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <typename T> class Scalar_ {
public:
    T val[4];
    Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0, T v1 = 0, T v2 = 0, T v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Scalar_<double> Scalar;

class TermCriteria {
public:
    enum Type { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
    int type, maxCount;
    double epsilon;
    TermCriteria() : type(0), maxCount(0), epsilon(0) {}
    TermCriteria(int t, int n, double e) : type(t), maxCount(n), epsilon(e) {}
};

template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8UC1 }; };
template <> struct DataType<int> { enum { type = CV_32SC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<double> { enum { type = CV_64FC1 }; };

// Dense 2-D array with shared ownership (copies are headers on the same buffer, like cv::Mat).
class Mat {
public:
    int rows, cols;
    size_t step;  // bytes per row
    uchar* data;

    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(0) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
    // header on caller memory (not owned)
    Mat(int r, int c, int type, void* ptr, size_t stp = 0) : rows(r), cols(c), data((uchar*)ptr), type_(type) {
        step = stp ? stp : (size_t)c * elemSize();
    }

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_ && buf_) return;
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        size_t n = step * (size_t)r;
        buf_ = std::shared_ptr<uchar>(new uchar[n ? n : 1](), std::default_delete<uchar[]>());  // value-init: zeros (C-1)
        data = buf_.get();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }

    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    size_t elemSize1() const { static const int s[7] = {1, 1, 2, 2, 4, 4, 8}; return (size_t)s[depth()]; }
    size_t elemSize() const { return elemSize1() * (size_t)channels(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return (size_t)rows * cols; }

    template <typename T> T& at(int i, int j) { return *(T*)(data + (size_t)i * step + (size_t)j * sizeof(T)); }
    template <typename T> const T& at(int i, int j) const { return *(const T*)(data + (size_t)i * step + (size_t)j * sizeof(T)); }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> T* ptr(int i = 0) { return (T*)(data + (size_t)i * step); }
    template <typename T> const T* ptr(int i = 0) const { return (const T*)(data + (size_t)i * step); }

    Mat operator()(const Rect& r) const {
        Mat m(*this);
        m.rows = r.height; m.cols = r.width;
        m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
        return m;
    }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int i = 0; i < rows; i++) std::memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }
    // Mat::convertTo (CylinderTag.cpp:80,101): dst = saturate_cast<rtype>(src * alpha + beta).
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;

protected:
    int type_;
    std::shared_ptr<uchar> buf_;
};

template <typename T> class Mat_ : public Mat {
public:
    Mat_() : Mat() { type_ = DataType<T>::type; }
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    Mat_(const Mat& m) : Mat(m) {
        if (!m.empty() && m.type() != DataType<T>::type) throw Exception("Mat_: type mismatch");
    }
    T& operator()(int i, int j) { return this->template at<T>(i, j); }
    const T& operator()(int i, int j) const { return this->template at<T>(i, j); }
    // element iteration (CylinderTag.cpp:30,58) -- continuous matrices only
    T* begin() { check(); return (T*)data; }
    T* end() { check(); return (T*)data + total(); }
    const T* begin() const { check(); return (const T*)data; }
    const T* end() const { check(); return (const T*)data + total(); }
private:
    void check() const { if (!empty() && !isContinuous()) throw Exception("Mat_ iteration: not continuous"); }
};
typedef Mat_<int> Mat1i;
typedef Mat_<float> Mat1f;
typedef Mat_<double> Mat1d;
typedef Mat_<uchar> Mat1b;

struct NoArray {};
inline NoArray noArray() { return NoArray(); }

enum ColorConversionCodes { COLOR_BGR2GRAY = 6, COLOR_RGB2GRAY = 7, COLOR_GRAY2BGR = 8, COLOR_GRAY2RGB = COLOR_GRAY2BGR };
enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum DistanceTypes { DIST_L1 = 1, DIST_L2 = 2, DIST_C = 3, DIST_L12 = 4, DIST_FAIR = 5, DIST_WELSCH = 6, DIST_HUBER = 7 };
enum ConnectedComponentsTypes { CC_STAT_LEFT = 0, CC_STAT_TOP = 1, CC_STAT_WIDTH = 2, CC_STAT_HEIGHT = 3, CC_STAT_AREA = 4 };
enum ConnectedComponentsAlgorithmsTypes { CCL_DEFAULT = -1, CCL_WU = 0, CCL_GRANA = 1, CCL_BOLELLI = 2, CCL_SAUF = 3, CCL_BBDT = 4, CCL_SPAGHETTI = 5 };
enum DecompTypes { DECOMP_LU = 0, DECOMP_SVD = 1 };
enum SolvePnPMethod { SOLVEPNP_ITERATIVE = 0, SOLVEPNP_EPNP = 1 };
enum LineTypes { FILLED = -1, LINE_4 = 4, LINE_8 = 8, LINE_AA = 16 };
enum HersheyFonts { FONT_HERSHEY_SIMPLEX = 0, FONT_ITALIC = 16 };

// ---- imgproc / core numerics ----------------------------------------------------------------------------------------
void cvtColor(const Mat& src, Mat& dst, int code);
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void minMaxLoc(const Mat& src, double* minVal, double* maxVal = nullptr, Point* minLoc = nullptr, Point* maxLoc = nullptr);
int connectedComponentsWithStats(const Mat& image, Mat& labels, Mat& stats, Mat& centroids, int connectivity, int ltype, int ccltype);
void fitLine(const std::vector<Point>& points, std::vector<float>& line, int distType, double param, double reps, double aeps);
void fitLine(const std::vector<Point2f>& points, std::vector<float>& line, int distType, double param, double reps, double aeps);
Scalar sum(const Mat& src);
double determinant(const Mat& m);
bool solve(const Mat& src1, const Mat& src2, Mat& dst, int flags = DECOMP_LU);
float fastAtan2(float y, float x);

// ---- calib3d (pose_estimation.cpp:96,109; CylinderTag.cpp:234) ------------------------------------------------------
bool solvePnP(const std::vector<Point3f>& objectPoints, const std::vector<Point2f>& imagePoints, const Mat& cameraMatrix,
              const Mat& distCoeffs, Mat& rvec, Mat& tvec, bool useExtrinsicGuess = false, int flags = SOLVEPNP_ITERATIVE);
void undistortPoints(const std::vector<Point2f>& src, std::vector<Point2f>& dst, const Mat& cameraMatrix, const Mat& distCoeffs,
                     NoArray R, const Mat& P);
void projectPoints(const std::vector<Point3f>& objectPoints, const Mat& rvec, const Mat& tvec, const Mat& cameraMatrix,
                   const Mat& distCoeffs, std::vector<Point2f>& imagePoints);

// ---- persistence (CylinderTag.cpp:193): the OpenCV-YAML subset of cameraParams.yml ----------------------------------
class FileNode {
public:
    FileNode() : ok(false), rows(0), cols(0), dt('f') {}
    bool ok;
    int rows, cols;
    char dt;
    std::vector<double> values;
};
void operator>>(const FileNode& n, Mat& m);
class FileStorage {
public:
    enum Mode { READ = 0, WRITE = 1 };
    FileStorage(const std::string& path, int mode);
    bool isOpened() const { return opened; }
    FileNode operator[](const std::string& key) const;
    FileNode operator[](const char* key) const { return (*this)[std::string(key)]; }
private:
    bool opened;
    std::vector<std::pair<std::string, FileNode>> nodes;
};

// ---- drawing / highgui: display is out of scope; these keep the call sites linkable and do nothing ------
inline void circle(Mat&, Point2f, int, const Scalar&, int = 1, int = LINE_8, int = 0) {}
inline void line(Mat&, Point2f, Point2f, const Scalar&, int = 1, int = LINE_8, int = 0) {}
inline void arrowedLine(Mat&, Point2f, Point2f, const Scalar&, int = 1, int = LINE_8, int = 0, double = 0.1) {}
inline void putText(Mat&, const std::string&, Point2f, int, double, Scalar, int = 1, int = LINE_8, bool = false) {}
void imshow(const std::string& winname, const Mat& img);  // remembers the last image per thread (shim_last_shown)
inline int waitKey(int = 0) { return -1; }
inline void destroyAllWindows() {}

}  // namespace cv

// ---- backend switch: route a primitive to the real OpenCV (cv2, through ctypes callbacks) instead of the restatement -
extern "C" {
typedef struct shim_backend {
    // 8-bit single-channel bicubic resize; return 0 on success
    int (*resize_cubic_u8)(const unsigned char* src, int sw, int sh, size_t sstep, unsigned char* dst, int dw, int dh, size_t dstep);
    // 8-connected BBDT labelling of a 0/non-0 image into int32 labels (w*h, dense); returns the label count incl. background
    int (*ccl_bbdt)(const unsigned char* img, int w, int h, size_t step, int* labels);
    // cv::fitLine on n float points (xy interleaved); dist = cv::DIST_*; writes vx, vy, x0, y0
    int (*fit_line)(const float* xy, int n, int dist, double param, double reps, double aeps, float* line4);
    float (*fast_atan2)(float y, float x);
    // cv::solvePnP(SOLVEPNP_EPNP); K is 3x3 float row-major, D has nd floats; rvec/tvec double[3]
    int (*solve_pnp_epnp)(const float* obj_xyz, const float* img_xy, int n, const float* K, const float* D, int nd, double* rvec, double* tvec);
    // cv::undistortPoints(src, K, D, noArray(), P = K)
    int (*undistort_points)(const float* xy, int n, const float* K, const float* D, int nd, float* out_xy);
    // cv::projectPoints
    int (*project_points)(const float* obj_xyz, int n, const double* rvec, const double* tvec, const float* K, const float* D, int nd, float* out_xy);
    // u8 -> f32 convertTo with scale (alpha as double), n elements
    int (*convert_u8_f32)(const unsigned char* src, int n, double alpha, float* dst);
} shim_backend;
// Installs (copy) the callbacks for the calling process; NULL members / NULL pointer = built-in restatement.
__attribute__((visibility("default"))) void shim_set_backend(const shim_backend* b);
}

#endif
