// oracle/cvshim -- TEST INFRASTRUCTURE.  A minimal stand-in for the slice of the OpenCV C++ API that
// the reference's src/ORBextractor.cc + include/ORBextractor.h use (includes at ORBextractor.cc:57-63,
// ORBextractor.h:26), so that the reference's own extractor source compiles UNMODIFIED from
// /root/reference into oracle/_ref/ (see oracle/Makefile).  All arithmetic lives in ../../cvprim.h,
// which is pinned bit-exactly against python cv2 4.13 golden vectors.  Nothing here is product code.
#pragma once
#include <cassert>
#include <cstdlib>
#include <memory>
#include <vector>
#include "../../../cvprim.h"

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8UC1 0
#define CV_PI 3.1415926535897932384626433832795

static inline int cvRound(double v) { return cvprim::round_half_even(v); }
static inline int cvRound(float v) { return cvprim::round_half_even_f(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { return cvprim::floor_i(v); }
static inline int cvCeil(double v) { return cvprim::ceil_i(v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {} };

struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

// 8-bit single-channel matrix header with shared ownership and ROI support.
class Mat {
public:
    int rows, cols;
    size_t step;
    uchar* data;
    std::shared_ptr<std::vector<uchar> > buf;

    Mat() : rows(0), cols(0), step(0), data(0) {}
    Mat(Size sz, int) { alloc(sz.height, sz.width); }
    Mat(int r, int c, int) { alloc(r, c); }
    Mat(int r, int c, int, void* ext, size_t st = 0) : rows(r), cols(c), step(st ? st : (size_t)c), data((uchar*)ext) {}
    void alloc(int r, int c) {
        rows = r; cols = c; step = (size_t)c;
        buf.reset(new std::vector<uchar>((size_t)r * c));
        data = buf->data();
    }
    void create(int r, int c, int) { if (r != rows || c != cols || !data) alloc(r, c); }
    void release() { rows = cols = 0; step = 0; data = 0; buf.reset(); }
    // Mat::zeros yields an expression; assigning it to a same-shape Mat (even an ROI header, as at
    // ORBextractor.cc:1037 where `descriptors` aliases a rowRange of the output) clears it IN PLACE.
    struct ZerosExpr { int r, c; operator Mat() const { Mat m(r, c, CV_8UC1); if (m.data) memset(m.data, 0, (size_t)r * c); return m; } };
    static ZerosExpr zeros(int r, int c, int) { ZerosExpr e; e.r = r; e.c = c; return e; }
    Mat& operator=(const ZerosExpr& e) {
        create(e.r, e.c, CV_8UC1);
        for (int y = 0; y < rows; y++) memset(data + (size_t)y * step, 0, cols);
        return *this;
    }
    int type() const { return CV_8UC1; }
    bool empty() const { return data == 0 || rows * cols == 0; }
    size_t step1() const { return step; }
    Mat operator()(const Rect& r) const { Mat m(*this); m.data = data + (size_t)r.y * step + r.x; m.rows = r.height; m.cols = r.width; return m; }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + a; m.cols = b - a; return m; }
    Mat clone() const {
        Mat m(rows, cols, CV_8UC1);
        for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols);
        return m;
    }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + x); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + x); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
};

// InputArray / OutputArray: thin references to a Mat.
class _InputArray {
public:
    const Mat* m;
    _InputArray() : m(0) {}
    _InputArray(const Mat& _m) : m(&_m) {}
    bool empty() const { return !m || m->empty(); }
    Mat getMat() const { return m ? *m : Mat(); }
};
class _OutputArray {
public:
    Mat* m;
    _OutputArray(Mat& _m) : m(&_m) {}
    void create(int r, int c, int t) const { m->create(r, c, t); }
    void release() const { m->release(); }
    Mat getMat() const { return *m; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

static inline float fastAtan2(float y, float x) { return cvprim::fast_atan2_deg(y, x); }

static inline void resize(const Mat& src, Mat& dst, Size dsize, double = 0, double = 0, int = INTER_LINEAR) {
    if (dst.rows != dsize.height || dst.cols != dsize.width || !dst.data) dst.alloc(dsize.height, dsize.width);
    cvprim::resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}
static inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int) {
    if (dst.rows != src.rows + top + bottom || dst.cols != src.cols + left + right || !dst.data)
        dst.alloc(src.rows + top + bottom, src.cols + left + right);
    cvprim::copy_make_border_reflect101(src.data, src.cols, src.rows, src.step, dst.data, dst.step, top, bottom, left, right);
}
static inline void GaussianBlur(const Mat& src, Mat& dst, Size ks, double sx, double sy, int border) {
    assert(ks.width == 7 && ks.height == 7 && sx == 2 && sy == 2 && border == BORDER_REFLECT_101);
    (void)ks; (void)sx; (void)sy; (void)border;
    if (dst.rows != src.rows || dst.cols != src.cols || !dst.data) dst.alloc(src.rows, src.cols);
    cvprim::gaussian_blur7_s2(src.data, src.cols, src.rows, src.step, dst.data, dst.step);
}
static inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nms = true) {
    assert(nms); (void)nms;
    std::vector<cvprim::FastKp> v;
    cvprim::fast9_16_nms(img.data, img.cols, img.rows, img.step, threshold, v);
    kps.clear();
    for (size_t i = 0; i < v.size(); i++) kps.push_back(KeyPoint((float)v[i].x, (float)v[i].y, 7.f, -1, (float)v[i].score));
}

struct KeyPointsFilter {
    // only referenced from the reference's dead ComputeKeyPointsOld (ORBextractor.cc:855-1032)
    static void retainBest(std::vector<KeyPoint>& k, int n) {
        if (n >= 0 && (int)k.size() > n) {
            std::stable_sort(k.begin(), k.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            k.resize(n);
        }
    }
};

}  // namespace cv
