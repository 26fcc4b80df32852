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
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
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
                else throw Exception(-1, "setTo: type", "setTo", __FILE__, __LINE__);
            }
        return *this;
    }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
private:
    int type_;
    std::shared_ptr<std::vector<uchar> > buf_;
};

class FileNode;
class FileStorage;
template <typename T> class Ptr : public std::shared_ptr<T> {      // cv::Ptr (markerlabeler.h:56,70)
public:
    Ptr() {}
    Ptr(T* p) : std::shared_ptr<T>(p) {}
    template <typename U> Ptr(const Ptr<U>& o) : std::shared_ptr<T>(o) {}
    bool empty() const { return !this->get(); }
};

enum { THRESH_BINARY = 0, THRESH_OTSU = 8, FONT_HERSHEY_COMPLEX = 3 };

// cv::threshold(src, dst, 125, 255, THRESH_BINARY | THRESH_OTSU) on CV_8UC1 (dictionary_based.cpp:1127), in place allowed
static inline double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type) {
    if (src.type() != CV_8UC1) throw Exception(-1, "threshold: CV_8UC1 only", "threshold", __FILE__, __LINE__);
    int level = (int)thresh;
    if (type & THRESH_OTSU) level = cvprim::otsu_level(src.data, src.cols, src.rows, src.step);
    Mat out = src.data == dst.data ? dst : Mat(src.rows, src.cols, CV_8UC1);
    for (int y = 0; y < src.rows; y++) for (int x = 0; x < src.cols; x++) out.at<uchar>(y, x) = src.at<uchar>(y, x) > level ? (uchar)maxval : 0;
    dst = out;
    return level;
}
static inline void cvtColor(const Mat&, Mat&, int) { throw Exception(-1, "cvtColor is not part of the stand-in (gray input only)", "cvtColor", __FILE__, __LINE__); }
static inline void putText(Mat&, const std::string&, Point, int, double, Scalar, int = 1) {}
static inline void circle(Mat&, Point2f, int, const Scalar&, int = 1) {}

}  // namespace cv
