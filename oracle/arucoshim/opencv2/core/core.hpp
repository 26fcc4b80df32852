// oracle/arucoshim/opencv2/core/core.hpp -- TEST INFRASTRUCTURE.  The part of cv:: that the reference's Thirdparty/aruco/aruco/dictionary.cpp,
// dictionary_based.cpp and markerlabeler.cpp (and the headers they pull: marker.h, markermap.h) need to compile UNMODIFIED into
// oracle/_ref/libref_dict.so: a reference-counted cv::Mat for CV_8UC1 / CV_32SC1 / CV_32FC1 with ROI views, Scalar, Range, points, Exception,
// and cv::threshold(THRESH_BINARY | THRESH_OTSU) on top of oracle/cvprim_aruco.h (itself pinned to cv2 golden vectors).  Drawing calls are no-ops
// (only used for the optional watermark of Dictionary::getMarkerImage_id).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <exception>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../../cvprim_aruco.h"

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32S 4
#define CV_32SC1 4
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
#define CV_BGR2GRAY 6
#define CV_MAJOR_VERSION 3
#define CV_Assert(x) do { if (!(x)) throw cv::Exception(-215, #x, __func__, __FILE__, __LINE__); } while (0)

namespace cv {

class Exception : public std::exception {
public:
    Exception() : code(0), line(0) {}
    Exception(int c, const std::string& e, const std::string& f, const std::string& fi, int l) : code(c), line(l), err(e), func(f), file(fi) { msg = file + ":" + func + ": " + err; }
    virtual ~Exception() throw() {}
    virtual const char* what() const throw() { return msg.c_str(); }
    int code, line; std::string err, func, file, msg;
};

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    template <typename S> Point_& operator*=(S k) { x = (T)(x * k); y = (T)(y * k); return *this; }      // saturate_cast of the product, as cv::Point_
    Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
    Point_& operator-=(const Point_& o) { x -= o.x; y -= o.y; return *this; }
    T dot(const Point_& o) const { return (T)(x * o.x + y * o.y); }
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
    Point3_& operator-=(const Point3_& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    Point3_& operator+=(const Point3_& o) { x += o.x; y += o.y; z += o.z; return *this; }
};
typedef Point3_<float> Point3f;
template <typename T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> static inline Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> static inline Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> static inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
template <typename T> static inline double norm(const Point3_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z); }
struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
};
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int a, int b, int w, int h) : x(a), y(b), width(w), height(h) {} };
struct Vec4i { int val[4]; };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {} };
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(0) {} };
struct TermCriteria { enum { COUNT = 1, MAX_ITER = 1, EPS = 2 }; int type, maxCount; double epsilon; TermCriteria(int t = 0, int c = 0, double e = 0) : type(t), maxCount(c), epsilon(e) {} };
struct Range { int start, end; Range() : start(0), end(0) {} Range(int s, int e) : start(s), end(e) {} };
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double operator[](int i) const { return val[i]; }
};

class Mat {
public:
    int rows, cols;
    size_t step;
    uchar* data;
    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(0) {}
    Mat(int r, int c, int t) : rows(0), cols(0), step(0), data(nullptr), type_(0) { create(r, c, t); }
    Mat(Size s, int t) : rows(0), cols(0), step(0), data(nullptr), type_(0) { create(s.height, s.width, t); }
    Mat(int r, int c, int t, const Scalar& sc) : rows(0), cols(0), step(0), data(nullptr), type_(0) { create(r, c, t); setTo(sc); }
    void create(Size sz, int t) { create(sz.height, sz.width, t); }
    Mat(const Mat& m, const Rect& r) : rows(r.height), cols(r.width), step(m.step), data(m.data + (size_t)r.y * m.step + (size_t)r.x * esz(m.type_)), type_(m.type_), buf_(m.buf_) {}
    Mat(int r, int c, int t, void* p, size_t st = 0) : rows(r), cols(c), step(st ? st : (size_t)c * esz(t)), data((uchar*)p), type_(t) {}
    static size_t esz(int t) { return t == CV_8UC1 ? 1 : t == CV_8UC3 ? 3 : t == CV_64F ? 8 : 4; }
    void create(int r, int c, int t) {
        if (data && r == rows && c == cols && t == type_) return;
        rows = r; cols = c; type_ = t; step = (size_t)c * esz(t);
        buf_ = std::make_shared<std::vector<uchar> >((size_t)r * step + 8, 0);
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int t) { Mat m(r, c, t); if (m.data) memset(m.data, 0, (size_t)r * m.step); return m; }
    int type() const { return type_; }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return !data || rows * cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& dst) const {                                  // into an existing same-shape view (e.g. an ROI) in place, else a fresh buffer
        if (!(dst.data && dst.rows == rows && dst.cols == cols && dst.type_ == type_)) { dst = Mat(); dst.create(rows, cols, type_); }
        for (int y = 0; y < rows; y++) memmove(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols * esz(type_));
    }
    Mat operator()(const Range& rr, const Range& cr) const {
        Mat m(*this);
        m.data = data + (size_t)rr.start * step + (size_t)cr.start * esz(type_); m.rows = rr.end - rr.start; m.cols = cr.end - cr.start;
        return m;
    }
    Mat& setTo(const Scalar& s) {
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) {
                if (type_ == CV_8UC1) at<uchar>(y, x) = (uchar)s.val[0];
                else if (type_ == CV_32SC1) at<int>(y, x) = (int)s.val[0];
                else if (type_ == CV_32FC1) at<float>(y, x) = (float)s.val[0];
                else if (type_ == CV_64F) at<double>(y, x) = s.val[0];
                else throw Exception(-1, "setTo: type", "setTo", __FILE__, __LINE__);
            }
        return *this;
    }
    // pose algebra (Marker::calculateExtrinsics, CameraParameters): declared so that marker.cpp compiles, never executed without camera parameters
    double get(int y, int x) const {
        return type_ == CV_8UC1 ? (double)at<uchar>(y, x) : type_ == CV_32SC1 ? (double)at<int>(y, x) : type_ == CV_32FC1 ? (double)at<float>(y, x) : at<double>(y, x);
    }
    void convertTo(Mat& dst, int t) const {                        // single-channel cast (CameraParameters::setParams, cameraparameters.cpp:84,90)
        Mat out(rows, cols, t);
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) {
                const double v = get(y, x);
                if (t == CV_32FC1) out.at<float>(y, x) = (float)v; else if (t == CV_64F) out.at<double>(y, x) = v;
                else if (t == CV_32SC1) out.at<int>(y, x) = (int)std::lrint(v); else if (t == CV_8UC1) out.at<uchar>(y, x) = (uchar)std::min(255.0, std::max(0.0, std::nearbyint(v)));
                else throw std::runtime_error("oracle/arucoshim: Mat::convertTo: type");
            }
        dst = out;
    }
    int depth() const { return type_; }
    static Mat eye(int, int, int) { throw std::runtime_error("oracle/arucoshim: Mat::eye is not on the reference's configured path"); }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * esz(type_); m.cols = b - a; return m; }
    Mat inv() const { throw std::runtime_error("oracle/arucoshim: Mat::inv is not on the reference's configured path"); }
    Mat t() const { throw std::runtime_error("oracle/arucoshim: Mat::t is not on the reference's configured path"); }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
private:
    int type_;
    std::shared_ptr<std::vector<uchar> > buf_;
};

typedef const Mat& InputArray;
typedef Mat& OutputArray;
#define CV_64FC1 6

// cv::FileStorage / cv::FileNode: named by the parameter save / load members, which are never called here - every member exists and refuses to work
struct NoFileStorage : std::runtime_error { NoFileStorage() : std::runtime_error("cv::FileStorage is not part of the stand-in") {} };
class FileNode {
public:
    enum { NONE = 0 };
    int type() const { throw NoFileStorage(); }
    FileNode operator[](const char*) const { throw NoFileStorage(); }
    FileNode operator[](const std::string&) const { throw NoFileStorage(); }
    FileNode operator[](int) const { throw NoFileStorage(); }
    size_t size() const { throw NoFileStorage(); }
    template <typename T> operator T() const { throw NoFileStorage(); }
};
template <typename T> void operator>>(const FileNode&, T&) { throw NoFileStorage(); }
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const char*, int) { throw NoFileStorage(); }
    FileStorage(const std::string&, int) { throw NoFileStorage(); }
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { throw NoFileStorage(); }
    FileNode operator[](const std::string&) const { throw NoFileStorage(); }
};
template <typename T> FileStorage& operator<<(FileStorage&, const T&) { throw NoFileStorage(); }
template <typename T> class Ptr : public std::shared_ptr<T> {      // cv::Ptr (markerlabeler.h:56,70)
public:
    Ptr() {}
    Ptr(T* p) : std::shared_ptr<T>(p) {}
    template <typename U> Ptr(const Ptr<U>& o) : std::shared_ptr<T>(o) {}
    bool empty() const { return !this->get(); }
};

enum { THRESH_BINARY = 0, THRESH_BINARY_INV = 1, THRESH_OTSU = 8, FONT_HERSHEY_COMPLEX = 3, FONT_HERSHEY_SIMPLEX = 0, ADAPTIVE_THRESH_MEAN_C = 0, RETR_LIST = 1, CHAIN_APPROX_NONE = 1,
       INTER_NEAREST = 0, INTER_LINEAR = 1, DECOMP_SVD = 1, MORPH_CROSS = 1 };
struct NotInStandIn : std::runtime_error { NotInStandIn(const char* w) : std::runtime_error(std::string("oracle/arucoshim: ") + w + " is not on the reference's configured path") {} };
struct NoArray {};
static inline NoArray noArray() { return NoArray(); }

// cv::threshold(src, dst, 125, 255, THRESH_BINARY | THRESH_OTSU) on CV_8UC1 (dictionary_based.cpp:1127), in place allowed
static inline double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type) {
    if (src.type() != CV_8UC1 || (type & 7) != THRESH_BINARY) throw NotInStandIn("threshold other than CV_8UC1 THRESH_BINARY[|OTSU]");
    int level = (int)thresh;
    if (type & THRESH_OTSU) level = cvprim::otsu_level(src.data, src.cols, src.rows, src.step);
    Mat out = src.data == dst.data ? dst : Mat(src.rows, src.cols, CV_8UC1);
    for (int y = 0; y < src.rows; y++) for (int x = 0; x < src.cols; x++) out.at<uchar>(y, x) = src.at<uchar>(y, x) > level ? (uchar)maxval : 0;
    dst = out;
    return level;
}
// cv::adaptiveThreshold(src, dst, 255, ADAPTIVE_THRESH_MEAN_C, THRESH_BINARY_INV, bs, C) (markerdetector_impl.cpp:2984)
static inline void adaptiveThreshold(const Mat& src, Mat& dst, double maxval, int method, int type, int bs, double Cc) {
    if (src.type() != CV_8UC1 || maxval != 255. || method != ADAPTIVE_THRESH_MEAN_C || type != THRESH_BINARY_INV) throw NotInStandIn("this adaptiveThreshold mode");
    dst.create(src.rows, src.cols, CV_8UC1);                        // keeps the buffer when the shape already matches, like cv::Mat::create
    cvprim::adaptive_threshold_mean_inv(src.data, src.cols, src.rows, src.step, dst.data, dst.step, bs, (int)Cc);
}
// cv::findContours(img, contours, noArray(), RETR_LIST, CHAIN_APPROX_NONE) (markerdetector_impl.cpp:3108)
static inline void findContours(const Mat& img, std::vector<std::vector<Point> >& contours, NoArray, int mode, int method) {
    if (img.type() != CV_8UC1 || mode != RETR_LIST || method != CHAIN_APPROX_NONE) throw NotInStandIn("this findContours mode");
    std::vector<std::vector<cvprim::Pt> > c;
    cvprim::find_contours_list_none(img.data, img.cols, img.rows, img.step, c);
    contours.resize(c.size());
    for (size_t i = 0; i < c.size(); i++) { contours[i].resize(c[i].size()); for (size_t k = 0; k < c[i].size(); k++) contours[i][k] = Point(c[i][k].x, c[i][k].y); }
}
static inline void approxPolyDP(const std::vector<Point>& curve, std::vector<Point>& approx, double eps, bool closed) {          // :3253
    if (!closed) throw NotInStandIn("open approxPolyDP");
    std::vector<cvprim::Pt> in(curve.size()), out;
    for (size_t i = 0; i < curve.size(); i++) { in[i].x = curve[i].x; in[i].y = curve[i].y; }
    cvprim::approx_poly_dp_closed(in, eps, out);
    approx.resize(out.size());
    for (size_t i = 0; i < out.size(); i++) approx[i] = Point(out[i].x, out[i].y);
}
static inline bool isContourConvex(const std::vector<Point>& p) {                                                              // :3292
    std::vector<cvprim::Pt> in(p.size());
    for (size_t i = 0; i < p.size(); i++) { in[i].x = p[i].x; in[i].y = p[i].y; }
    return cvprim::is_contour_convex(in);
}
// cv::resize(src, dst, dsize): INTER_LINEAR; the detector's pyramid halves exactly (:1386-1466), where OpenCV averages 2 x 2 blocks
static inline void resize(const Mat& src, Mat& dst, Size dsize, double = 0, double = 0, int interp = INTER_LINEAR) {
    if (src.type() != CV_8UC1 || interp != INTER_LINEAR) throw NotInStandIn("this resize mode");
    Mat out(dsize.height, dsize.width, CV_8UC1);
    if (dsize.width * 2 == src.cols && dsize.height * 2 == src.rows) cvprim::resize_half(src.data, src.cols, src.rows, src.step, out.data, dsize.width, dsize.height, out.step);
    else cvprim::resize_linear_u8(src.data, src.cols, src.rows, src.step, out.data, dsize.width, dsize.height, out.step);
    dst = out;
}
static inline Mat getPerspectiveTransform(const Point2f src[], const Point2f dst[]) {                                           // :11079
    float s8[8], d8[8];
    for (int i = 0; i < 4; i++) { s8[2 * i] = src[i].x; s8[2 * i + 1] = src[i].y; d8[2 * i] = dst[i].x; d8[2 * i + 1] = dst[i].y; }
    Mat M(3, 3, CV_64F);
    cvprim::get_perspective_transform(s8, d8, M.ptr<double>(0));
    return M;
}
static inline void warpPerspective(const Mat& src, Mat& dst, const Mat& M, Size dsize, int flags) {                              // :11092
    if (src.type() != CV_8UC1 || flags != INTER_LINEAR || M.type() != CV_64F) throw NotInStandIn("this warpPerspective mode");
    Mat out(dsize.height, dsize.width, CV_8UC1);
    cvprim::warp_perspective_linear(src.data, src.cols, src.rows, src.step, out.data, dsize.width, dsize.height, out.step, M.ptr<double>(0));
    dst = out;
}
static inline int countNonZero(const Mat& m) { int n = 0; for (int y = 0; y < m.rows; y++) for (int x = 0; x < m.cols; x++) n += m.get(y, x) != 0; return n; }
static inline void minMaxIdx(const Mat& m, double* mn, double* mx) {
    double a = 255, b = 0;
    for (int y = 0; y < m.rows; y++) for (int x = 0; x < m.cols; x++) { const double v = m.at<uchar>(y, x); if (v < a) a = v; if (v > b) b = v; }
    if (mn) *mn = a; if (mx) *mx = b;
}
// cv::solve(A, B, X, DECOMP_SVD) on CV_32FC1 (:11669, 11841, 12049)
static inline bool solve(const Mat& A, const Mat& B, Mat& X, int flags) {
    if (A.type() != CV_32FC1 || B.type() != CV_32FC1 || flags != DECOMP_SVD || B.cols != 1) throw NotInStandIn("this solve mode");
    std::vector<float> a((size_t)A.rows * A.cols), b(A.rows);
    for (int y = 0; y < A.rows; y++) { for (int x = 0; x < A.cols; x++) a[(size_t)y * A.cols + x] = A.at<float>(y, x); b[y] = B.at<float>(y, 0); }
    X = Mat(A.cols, 1, CV_32FC1);
    cvprim::solve_svd_f32(a.data(), b.data(), A.rows, A.cols, X.ptr<float>(0));
    return true;
}
// compiled but never executed with the reference's settings (src/Frame.cc:129-139): gray input, THRES_ADAPTIVE, CORNER_LINES, no camera parameters
static inline void cvtColor(const Mat&, Mat&, int) { throw NotInStandIn("cvtColor"); }
static inline void Rodrigues(const Mat&, Mat&) { throw NotInStandIn("Rodrigues"); }
static inline void Rodrigues(const Mat&, Mat&, Mat&) { throw NotInStandIn("Rodrigues"); }
static inline Mat operator*(const Mat&, const Mat&) { throw NotInStandIn("Mat product"); }
static inline Mat getStructuringElement(int, Size, Point) { throw NotInStandIn("getStructuringElement"); }
static inline void erode(const Mat&, Mat&, const Mat&) { throw NotInStandIn("erode"); }
static inline void bitwise_xor(const Mat&, const Mat&, Mat&) { throw NotInStandIn("bitwise_xor"); }
static inline void cornerSubPix(const Mat&, std::vector<Point2f>&, Size, Size, TermCriteria) { throw NotInStandIn("cornerSubPix"); }
static inline void undistortPoints(const std::vector<Point2f>&, std::vector<Point2f>&, const Mat&, const Mat&, const Mat&, const Mat&) { throw NotInStandIn("undistortPoints"); }
static inline void projectPoints(const std::vector<Point3f>&, const Mat&, const Mat&, const Mat&, const Mat&, std::vector<Point2f>&) { throw NotInStandIn("projectPoints"); }
static inline void putText(Mat&, const std::string&, Point, int, double, Scalar, int = 1) {}
static inline void circle(Mat&, Point2f, int, const Scalar&, int = 1) {}
static inline void line(Mat&, Point2f, Point2f, const Scalar&, int = 1) {}
static inline void rectangle(Mat&, Point2f, Point2f, const Scalar&, int = 1) {}
static inline void drawContours(Mat&, const std::vector<std::vector<Point> >&, int, const Scalar&) {}

}  // namespace cv
