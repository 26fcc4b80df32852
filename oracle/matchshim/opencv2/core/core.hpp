// oracle/matchshim -- TEST INFRASTRUCTURE.  A minimal stand-in for the slice of the OpenCV C++ API that the reference's
// src/ORBmatcher.cc uses (includes at ORBmatcher.cc:25-26): cv::KeyPoint / Point2f and a cv::Mat that is either a row-major CV_8U
// descriptor table or a small CV_32F matrix with the handful of algebra operators the projection code applies.  With it (and
// slam_types.h) the reference's own matcher source compiles UNMODIFIED from /root/reference into oracle/_ref/libref_match.so (see
// oracle/Makefile).  The float algebra is only used for the projections the reference computes BEFORE it queries the frame grid; the
// wrapper records those queries, so nothing here has to round like OpenCV's gemm.  Nothing here is product code.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
};
typedef Point_<float> Point2f;

struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

class Mat {
public:
    int rows, cols, tp;
    size_t step;                       // bytes between rows
    uchar* data;
    std::shared_ptr<std::vector<uchar> > buf;

    Mat() : rows(0), cols(0), tp(CV_8U), step(0), data(0) {}
    Mat(int r, int c, int t) { alloc(r, c, t); }
    Mat(int r, int c, int t, void* ext) : rows(r), cols(c), tp(t), step((size_t)c * esz(t)), data((uchar*)ext) {}
    static size_t esz(int t) { return t == CV_32F ? 4 : 1; }
    void alloc(int r, int c, int t) {
        rows = r; cols = c; tp = t; step = (size_t)c * esz(t);
        buf.reset(new std::vector<uchar>((size_t)r * step, 0));
        data = buf->data();
    }
    // create / release / zeros: used by the vendored DBoW2 (FORB.cpp), which shares this stand-in through oracle/vocshim
    void create(int r, int c, int t) { if (!(data && r == rows && c == cols && t == tp)) alloc(r, c, t); }
    void release() { rows = cols = 0; step = 0; data = 0; buf.reset(); }
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
    void copyTo(Mat& dst) const;                               // src/MapPoint.cc:39 Pos.copyTo(mWorldPos)
    int type() const { return tp; }
    bool empty() const { return data == 0 || rows * cols == 0; }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * esz(tp); m.cols = b - a; return m; }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat col(int x) const { return colRange(x, x + 1); }
    Mat clone() const {
        Mat m(rows, cols, tp);
        for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * esz(tp));
        return m;
    }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    // single index: element i of a row or column vector (cv::Mat::at(int i0))
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    float f(int y, int x) const { return at<float>(y, x); }
    Mat t() const {
        Mat m(cols, rows, CV_32F);
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) m.at<float>(x, y) = f(y, x);
        return m;
    }
    double dot(const Mat& o) const {
        double s = 0;
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) s += (double)f(y, x) * (double)o.f(y, x);
        return s;
    }
};

inline void Mat::copyTo(Mat& dst) const { dst = clone(); }

static inline Mat operator*(const Mat& a, const Mat& b) {
    assert(a.cols == b.rows);
    Mat m(a.rows, b.cols, CV_32F);
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < b.cols; x++) {
            double s = 0;
            for (int k = 0; k < a.cols; k++) s += (double)a.f(y, k) * (double)b.f(k, x);
            m.at<float>(y, x) = (float)s;
        }
    return m;
}
template <typename F> static inline Mat map2(const Mat& a, const Mat& b, F fn) {
    assert(a.rows == b.rows && a.cols == b.cols);
    Mat m(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) m.at<float>(y, x) = fn(a.f(y, x), b.f(y, x));
    return m;
}
template <typename F> static inline Mat map1(const Mat& a, F fn) {
    Mat m(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) m.at<float>(y, x) = fn(a.f(y, x));
    return m;
}
static inline Mat operator+(const Mat& a, const Mat& b) { return map2(a, b, [](float p, float q) { return p + q; }); }
static inline Mat operator-(const Mat& a, const Mat& b) { return map2(a, b, [](float p, float q) { return p - q; }); }
static inline Mat operator-(const Mat& a) { return map1(a, [](float p) { return -p; }); }
static inline Mat operator*(double s, const Mat& a) { return map1(a, [s](float p) { return (float)(s * p); }); }
static inline Mat operator*(const Mat& a, double s) { return s * a; }
static inline Mat operator/(const Mat& a, double s) { return map1(a, [s](float p) { return (float)(p / s); }); }
static inline double norm(const Mat& a) { return std::sqrt(a.dot(a)); }

}  // namespace cv
