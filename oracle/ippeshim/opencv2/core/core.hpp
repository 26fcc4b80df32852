// oracle/ippeshim -- TEST INFRASTRUCTURE.  A stand-in for the slice of OpenCV that the reference's Thirdparty/aruco/aruco/ippe.cpp uses, so that the pose
// solver (IPPE::PoseSolver, aruco::solvePnP: ippe.cpp:72-1169) compiles UNMODIFIED from /root/reference into oracle/_ref/libref_ippe.so (oracle/Makefile).
// What is the reference's own: every statement of the solver (canonical object points, Harker-O'Leary homography, the two rotations and translations,
// reprojection-error sort, rotation vector).  What is restated here: cv::Mat as a typed view with shared storage, Input/OutputArray, and the five OpenCV
// primitives the solver calls - undistortPoints (5 fixed-point iterations in double, float result: the restatement pinned bit-exact to cv2 4.13 in
// tests/golden/frame.npz), projectPoints (pinhole + k1 k2 p1 p2 k3, double), Rodrigues (vector -> matrix), eigen (cyclic Jacobi, descending eigenvalues,
// eigenvectors in rows) and transpose.  Nothing here is product code.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_64FC2 CV_MAKETYPE(CV_64F, 2)
#define CV_64FC3 CV_MAKETYPE(CV_64F, 3)

namespace cv {

template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
    Vec(T a, T b) { val[0] = a; val[1] = b; }
    Vec(T a, T b, T c) { val[0] = a; val[1] = b; val[2] = c; }
    T& operator()(int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<double, 2> Vec2d; typedef Vec<double, 3> Vec3d; typedef Vec<float, 2> Vec2f; typedef Vec<float, 3> Vec3f;
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point_<float> Point2f; typedef Point_<double> Point2d; typedef Point3_<float> Point3f; typedef Point3_<double> Point3d;
struct Rect { int x, y, width, height; Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };

class _InputArray;
class _OutputArray;
struct MatExpr;

class Mat {
public:
    // cv::Mat = cv::MatExpr evaluates INTO the existing buffer when shape and type already fit (Mat::create is a no-op then), so a local header that
    // shares storage with the caller's matrix updates the caller too (ippe.cpp relies on it: `H = H / H.at<double>(2, 2)` on _H.getMat())
    inline Mat& operator=(const MatExpr& e);
    Mat(const Mat&) = default;
    Mat& operator=(const Mat&) = default;
    int rows, cols, tp;
    size_t step;
    uchar* data;
    std::shared_ptr<std::vector<uchar> > buf;
    Mat() : rows(0), cols(0), tp(0), step(0), data(0) {}
    Mat(int r, int c, int t) { alloc(r, c, t); }
    Mat(int r, int c, int t, void* ext) : rows(r), cols(c), tp(t), step((size_t)c * esz(t)), data((uchar*)ext) {}
    Mat(const Mat& m, const Rect& roi) : rows(roi.height), cols(roi.width), tp(m.tp), step(m.step), data(m.data + (size_t)roi.y * m.step + (size_t)roi.x * esz(m.tp)), buf(m.buf) {}
    explicit Mat(const Point3d& p) { alloc(3, 1, CV_64FC1); at<double>(0) = p.x; at<double>(1) = p.y; at<double>(2) = p.z; }
    static size_t esz1(int t) { const int d = t & 7; return d == CV_64F ? 8 : d == CV_32F ? 4 : 1; }
    static int cn(int t) { return (t >> 3) + 1; }
    static size_t esz(int t) { return esz1(t) * cn(t); }
    void alloc(int r, int c, int t) {
        rows = r; cols = c; tp = t; step = (size_t)c * esz(t);
        buf.reset(new std::vector<uchar>((size_t)r * step + 8, 0));
        data = buf->data();
    }
    void create(int r, int c, int t) { if (!(data && r == rows && c == cols && t == tp)) alloc(r, c, t); }
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
    static Mat eye(int r, int c, int t) { Mat m(r, c, t); for (int i = 0; i < r && i < c; i++) m.set(i, i, 1.0); return m; }
    int type() const { return tp; }
    int depth() const { return tp & 7; }
    int channels() const { return cn(tp); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == 0 || rows * cols == 0; }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * esz(tp); m.cols = b - a; return m; }
    // element (y, x, channel) as double, whatever the depth
    double get(int y, int x, int c = 0) const {
        const uchar* p = data + (size_t)y * step + (size_t)x * esz(tp) + (size_t)c * esz1(tp);
        return depth() == CV_64F ? *(const double*)p : depth() == CV_32F ? (double)*(const float*)p : (double)*p;
    }
    void set(int y, int x, double v, int c = 0) {
        uchar* p = data + (size_t)y * step + (size_t)x * esz(tp) + (size_t)c * esz1(tp);
        if (depth() == CV_64F) *(double*)p = v; else if (depth() == CV_32F) *(float*)p = (float)v; else *p = (uchar)v;
    }
    Mat clone() const { Mat m(rows, cols, tp); for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * esz(tp)); return m; }
    void copyTo(Mat& dst) const {                                     // into an existing view of the same shape (ROI), else a fresh matrix
        if (!(dst.data && dst.rows == rows && dst.cols == cols && dst.tp == tp)) dst.alloc(rows, cols, tp);
        for (int y = 0; y < rows; y++) memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols * esz(tp));
    }
    inline void copyTo(const _OutputArray& dst) const;
    void convertTo(Mat& dst, int t) const {
        const int nt = CV_MAKETYPE(t & 7, channels());
        Mat out(rows, cols, nt);
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) for (int c = 0; c < channels(); c++) out.set(y, x, get(y, x, c), c);
        dst = out;
    }
    inline void convertTo(const _OutputArray& dst, int t) const;
    Mat& setTo(double v) { for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) for (int c = 0; c < channels(); c++) set(y, x, v, c); return *this; }
    inline MatExpr t() const;
};
struct MatExpr : Mat { explicit MatExpr(const Mat& m) : Mat(m) {} };
inline Mat& Mat::operator=(const MatExpr& e) {
    if (data && rows == e.rows && cols == e.cols && tp == e.tp) { for (int y = 0; y < rows; y++) memcpy(data + (size_t)y * step, e.data + (size_t)y * e.step, (size_t)cols * esz(tp)); }
    else *this = static_cast<const Mat&>(e);
    return *this;
}
inline MatExpr Mat::t() const { Mat m(cols, rows, tp); for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) m.set(x, y, get(y, x)); return MatExpr(m); }

// matrix product / sums in the operands' depth (double accumulation, one rounding into the result depth like cv::gemm)
static inline MatExpr operator*(const Mat& a, const Mat& b) {
    assert(a.cols == b.rows && a.channels() == 1 && b.channels() == 1);
    Mat m(a.rows, b.cols, a.depth() == CV_64F || b.depth() == CV_64F ? CV_64FC1 : CV_32FC1);
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < b.cols; x++) { double s = 0; for (int k = 0; k < a.cols; k++) s += a.get(y, k) * b.get(k, x); m.set(y, x, s); }
    return MatExpr(m);
}
template <typename F> static inline MatExpr zip(const Mat& a, const Mat& b, F f) {
    assert(a.rows == b.rows && a.cols == b.cols);
    Mat m(a.rows, a.cols, a.depth() == CV_64F || b.depth() == CV_64F ? CV_64FC1 : CV_32FC1);
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) m.set(y, x, f(a.get(y, x), b.get(y, x)));
    return MatExpr(m);
}
static inline MatExpr operator+(const Mat& a, const Mat& b) { return zip(a, b, [](double p, double q) { return p + q; }); }
static inline MatExpr operator-(const Mat& a, const Mat& b) { return zip(a, b, [](double p, double q) { return p - q; }); }
static inline MatExpr operator*(const Mat& a, double s) { Mat m(a.rows, a.cols, a.tp); for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) m.set(y, x, a.get(y, x) * s); return MatExpr(m); }
static inline MatExpr operator*(double s, const Mat& a) { return a * s; }
static inline MatExpr operator/(const Mat& a, double s) { Mat m(a.rows, a.cols, a.tp); for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) m.set(y, x, a.get(y, x) / s); return MatExpr(m); }
static inline MatExpr operator-(const Mat& a) { return a * -1.0; }

// cv::InputArray / OutputArray over a Mat or a std::vector of points (the two things ippe.cpp passes)
class _InputArray {
public:
    Mat m; bool none;
    _InputArray() : none(true) {}
    _InputArray(const Mat& x) : m(x), none(false) {}
    _InputArray(const std::vector<Point2f>& v) : m(1, (int)v.size(), CV_32FC2, (void*)v.data()), none(false) {}
    _InputArray(const std::vector<Point3f>& v) : m(1, (int)v.size(), CV_32FC3, (void*)v.data()), none(false) {}
    _InputArray(const std::vector<Point2d>& v) : m(1, (int)v.size(), CV_64FC2, (void*)v.data()), none(false) {}
    _InputArray(const std::vector<Point3d>& v) : m(1, (int)v.size(), CV_64FC3, (void*)v.data()), none(false) {}
    Mat getMat() const { return m; }
    int rows() const { return m.rows; }
    int cols() const { return m.cols; }
    int type() const { return m.type(); }
    int depth() const { return m.depth(); }
    bool empty() const { return none || m.empty(); }
    size_t total() const { return m.total(); }
    inline void copyTo(const _OutputArray& dst) const;
};
class _OutputArray {
public:
    Mat* pm;
    _OutputArray() : pm(0) {}
    _OutputArray(Mat& x) : pm(&x) {}
    _OutputArray(const Mat& x) : own(x), pm(&own) {}                   // a temporary ROI (R.copyTo(MRot.colRange(0, 3).rowRange(0, 3))): writes go through to the shared storage
    void create(int r, int c, int t) const { pm->create(r, c, t); }
    void setTo(double v) const { pm->setTo(v); }
    Mat getMat() const { return *pm; }
    bool needed() const { return pm != 0; }
private:
    mutable Mat own;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
static inline _InputArray noArray() { return _InputArray(); }
inline void Mat::copyTo(const _OutputArray& dst) const { copyTo(*dst.pm); }
inline void Mat::convertTo(const _OutputArray& dst, int t) const { convertTo(*dst.pm, t); }
inline void _InputArray::copyTo(const _OutputArray& dst) const { m.copyTo(*dst.pm); }

static inline void transpose(InputArray a, OutputArray d) { a.getMat().t().copyTo(*d.pm); }
static inline double determinant(InputArray a_) {
    const Mat a = a_.getMat();
    assert(a.rows == 3 && a.cols == 3);
    return a.get(0, 0) * (a.get(1, 1) * a.get(2, 2) - a.get(1, 2) * a.get(2, 1)) - a.get(0, 1) * (a.get(1, 0) * a.get(2, 2) - a.get(1, 2) * a.get(2, 0)) +
           a.get(0, 2) * (a.get(1, 0) * a.get(2, 1) - a.get(1, 1) * a.get(2, 0));
}

// symmetric eigen-decomposition: eigenvalues descending (n x 1), eigenvectors as ROWS, cyclic Jacobi in double
static inline bool eigen(InputArray src_, OutputArray evals, OutputArray evecs) {
    const Mat src = src_.getMat();
    const int n = src.rows;
    std::vector<double> A((size_t)n * n), V((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) A[i * n + j] = src.get(i, j); V[i * n + i] = 1.0; }
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0;
        for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
            if (std::fabs(A[p * n + q]) < 1e-300) continue;
            const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * A[p * n + q]);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0)), c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < n; k++) { const double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
            for (int k = 0; k < n; k++) { const double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
            for (int k = 0; k < n; k++) { const double vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq; }
        }
    }
    std::vector<int> ord(n);
    for (int i = 0; i < n; i++) ord[i] = i;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (A[ord[j] * n + ord[j]] > A[ord[i] * n + ord[i]]) std::swap(ord[i], ord[j]);
    evals.create(n, 1, CV_64FC1); evecs.create(n, n, CV_64FC1);
    Mat w = evals.getMat(), v = evecs.getMat();
    for (int i = 0; i < n; i++) { w.at<double>(i, 0) = A[ord[i] * n + ord[i]]; for (int k = 0; k < n; k++) v.at<double>(i, k) = V[k * n + ord[i]]; }
    return true;
}

// thin 3 x 3 SVD through the eigen-decomposition of A^T A (only computeObjextSpaceRSvD uses it: object points off the z = 0 plane, never a marker)
class SVD {
public:
    Mat u, w, vt;
    SVD() {}
    static void compute(InputArray a, OutputArray w_, OutputArray u_, OutputArray vt_, int = 0) { SVD s; s(a); s.w.copyTo(*w_.pm); s.u.copyTo(*u_.pm); s.vt.copyTo(*vt_.pm); }
    SVD& operator()(InputArray a_, int = 0) {
        const Mat a = a_.getMat();
        Mat ata = a.t() * a, ev, evec;
        eigen(ata, ev, evec);
        const int n = a.cols;
        w = Mat(n, 1, CV_64FC1); vt = evec.clone(); u = Mat(a.rows, n, CV_64FC1);
        for (int i = 0; i < n; i++) {
            const double s = std::sqrt(std::max(ev.at<double>(i, 0), 0.0));
            w.at<double>(i, 0) = s;
            for (int r = 0; r < a.rows; r++) { double acc = 0; for (int k = 0; k < n; k++) acc += a.get(r, k) * vt.get(i, k); u.at<double>(r, i) = s > 0 ? acc / s : 0.0; }
        }
        return *this;
    }
};

// rotation vector (3 x 1 or 1 x 3) -> rotation matrix, written into dst (possibly a float ROI: getRTMatrix, ippe.cpp:16-60); matrix -> vector is not needed
static inline void Rodrigues(InputArray src_, OutputArray dst) {
    const Mat src = src_.getMat();
    if (src.total() != 3) throw std::runtime_error("ippeshim: Rodrigues(matrix) is not part of the stand-in");
    const double rx = src.get(src.rows == 1 ? 0 : 0, 0), ry = src.rows == 1 ? src.get(0, 1) : src.get(1, 0), rz = src.rows == 1 ? src.get(0, 2) : src.get(2, 0);
    const double theta = std::sqrt(rx * rx + ry * ry + rz * rz);
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (theta > 2.220446049250313e-16) {
        const double c = std::cos(theta), s = std::sin(theta), c1 = 1. - c, itheta = 1. / theta, x = rx * itheta, y = ry * itheta, z = rz * itheta;
        const double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z}, rx_[9] = {0, -z, y, z, 0, -x, -y, x, 0};
        for (int k = 0; k < 9; k++) R[k] = c * (k % 4 == 0 ? 1.0 : 0.0) + c1 * rrt[k] + s * rx_[k];
    }
    Mat d = dst.getMat();
    if (!(d.rows == 3 && d.cols == 3)) { dst.create(3, 3, src.depth() == CV_32F ? CV_32FC1 : CV_64FC1); d = dst.getMat(); }
    for (int i = 0; i < 9; i++) d.set(i / 3, i % 3, R[i]);
}

// cv::undistortPoints(src, dst, K, D) without R / P: normalised coordinates, 5 iterations (OpenCV's default criteria), double arithmetic, result in
// the source's depth - the restatement oracle/frame_oracle.cpp pins bit-exact against cv2 4.13
static inline void undistortPoints(InputArray src_, OutputArray dst, InputArray K_, InputArray D_) {
    const Mat src = src_.getMat(), K = K_.getMat();
    double k[5] = {0, 0, 0, 0, 0};
    if (!D_.empty()) { const Mat D = D_.getMat(); const int nd = (int)D.total(); for (int i = 0; i < nd && i < 5; i++) k[i] = D.rows == 1 ? D.get(0, i) : D.get(i, 0); }
    const double fx = K.get(0, 0), fy = K.get(1, 1), cx = K.get(0, 2), cy = K.get(1, 2), ifx = 1. / fx, ify = 1. / fy;
    const int n = (int)src.total();
    dst.create(src.rows, src.cols, src.type());
    Mat d = dst.getMat();
    for (int i = 0; i < n; i++) {
        const int yy = src.rows == 1 ? 0 : i, xx = src.rows == 1 ? i : 0;
        double x = (src.get(yy, xx, 0) - cx) * ifx, y = (src.get(yy, xx, 1) - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y, icdist = 1. / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x), deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
            x = (x0 - deltaX) * icdist; y = (y0 - deltaY) * icdist;
        }
        d.set(yy, xx, x, 0); d.set(yy, xx, y, 1);
    }
}

// cv::projectPoints: X = R p + t, pinhole with k1 k2 p1 p2 k3, all in double; the result has the depth of the object points
static inline void projectPoints(InputArray obj_, InputArray rvec, InputArray tvec_, InputArray K_, InputArray D_, OutputArray img) {
    const Mat obj = obj_.getMat(), K = K_.getMat(), tv = tvec_.getMat();
    Mat R;
    Rodrigues(rvec, R);
    double k[5] = {0, 0, 0, 0, 0};
    if (!D_.empty()) { const Mat D = D_.getMat(); const int nd = (int)D.total(); for (int i = 0; i < nd && i < 5; i++) k[i] = D.rows == 1 ? D.get(0, i) : D.get(i, 0); }
    const double fx = K.get(0, 0), fy = K.get(1, 1), cx = K.get(0, 2), cy = K.get(1, 2);
    const double t[3] = {tv.rows == 1 ? tv.get(0, 0) : tv.get(0, 0), tv.rows == 1 ? tv.get(0, 1) : tv.get(1, 0), tv.rows == 1 ? tv.get(0, 2) : tv.get(2, 0)};
    const int n = (int)obj.total();
    img.create(obj.rows, obj.cols, CV_MAKETYPE(obj.depth(), 2));
    Mat d = img.getMat();
    for (int i = 0; i < n; i++) {
        const int yy = obj.rows == 1 ? 0 : i, xx = obj.rows == 1 ? i : 0;
        const double X = obj.get(yy, xx, 0), Y = obj.get(yy, xx, 1), Z = obj.get(yy, xx, 2);
        const double x = R.get(0, 0) * X + R.get(0, 1) * Y + R.get(0, 2) * Z + t[0], y = R.get(1, 0) * X + R.get(1, 1) * Y + R.get(1, 2) * Z + t[1];
        double z = R.get(2, 0) * X + R.get(2, 1) * Y + R.get(2, 2) * Z + t[2];
        z = z ? 1. / z : 1;
        const double xn = x * z, yn = y * z, r2 = xn * xn + yn * yn, r4 = r2 * r2, r6 = r4 * r2;
        const double a1 = 2 * xn * yn, a2 = r2 + 2 * xn * xn, a3 = r2 + 2 * yn * yn, cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6;
        d.set(yy, xx, (xn * cdist + k[2] * a1 + k[3] * a2) * fx + cx, 0);
        d.set(yy, xx, (yn * cdist + k[2] * a3 + k[3] * a1) * fy + cy, 1);
    }
}

}  // namespace cv
